"""JointsMSE loss with the reference's API (src/loss/mse.py:14-44): one fused sm_100a kernel computes
the loss over ALL stacks and, when autograd needs it, the gradient w.r.t. every stack's heat map in
the same pass (the reference launches 3 kernels per joint per stack)."""
from __future__ import absolute_import
from __future__ import division
from __future__ import print_function

import torch
import torch.nn as nn

from hgb200 import ops

__all__ = ['MSELoss', 'JointsMSELossOnTheFly']


class _FusedJMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, target_weight, mu, sigma, *outputs):
        need_grad = any(o.requires_grad for o in outputs)
        loss, grads = ops.jmse_loss([o.detach() for o in outputs], target, target_weight, want_grad=need_grad,
                                    mu=mu, sigma=sigma)
        ctx.grads = grads
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.grads
        ctx.grads = None
        return (None, None, None, None) + tuple(g * grad_out for g in grads)


class MSELoss(nn.Module):
    def __init__(self, use_target_weight):
        super(MSELoss, self).__init__()
        self.use_target_weight = use_target_weight

    def forward(self, outputs, target, target_weight):
        """outputs: list of [B,J,h,w]; target [B,J,h,w]; target_weight [B,J,1] -> 0-dim tensor."""
        outputs = list(outputs)
        dev = outputs[0].device
        if dev.type != 'cuda':
            raise RuntimeError("MSELoss (B200 build) runs on CUDA tensors only: there is no CPU fallback")
        target = target.to(dev, torch.float32).contiguous()
        tw = target_weight.to(dev, torch.float32).contiguous() if self.use_target_weight else None
        return _FusedJMSE.apply(target, tw, None, 1, *outputs)


class JointsMSELossOnTheFly(nn.Module):
    """Same loss, but the Gaussian target (src/datasets/common.py:197-248) is regenerated inside the
    kernel from joint coordinates, so no [B,J,h,w] target tensor is ever materialised or copied H2D."""

    def __init__(self, image_size, heatmap_size, sigma=1):
        super().__init__()
        self.image_size, self.heatmap_size, self.sigma = tuple(image_size), tuple(heatmap_size), sigma

    def forward(self, outputs, joints, joints_vis):
        """joints / joints_vis: float64 [B,J,3] in input-pixel coordinates (what JointsDataset feeds
        generate_target)."""
        outputs = list(outputs)
        dev = outputs[0].device
        joints = joints.to(dev, torch.float64).contiguous()
        joints_vis = joints_vis.to(dev, torch.float64).contiguous()
        mu, wt = ops.joint_centers(joints, joints_vis, self.heatmap_size, self.image_size, self.sigma)
        return _FusedJMSE.apply(None, wt, mu, self.sigma, *outputs)
