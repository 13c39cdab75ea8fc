"""fp32 CPU restatement of one reference training step (forward in train mode, JointsMSE loss,
backward, RMSprop update).

Test infrastructure only (see oracle/__init__.py).  The arithmetic of the reference's step lives in
PyTorch (unpinned; torch 2.11 here): autograd over Conv2d / BatchNorm2d(train) / ReLU / max_pool2d /
interpolate(nearest) / MSELoss, and torch.optim.RMSprop.  This restatement drives the functional
forward of oracle/hourglass_oracle.py with train-mode BatchNorm through the same autograd engine and
restates the optimizer update in closed form.  Pinned against the live reference by
tests/golden/train_s2_j16_64.npz (oracle/make_golden.py:train_golden).

Reference followed:
  * Trainer._train_epoch step          src/runner/trainer.py:82-99
  * optimizer construction             src/runner/trainer.py:39-41   (RMSprop, torch defaults)
  * adjust_learning_rate               src/runner/trainer.py:15-21
  * MSELoss                            src/loss/mse.py:14-44
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import hourglass_oracle as H

RMSPROP_ALPHA = 0.99     # torch.optim.RMSprop defaults (trainer.py:39-41 passes lr, momentum=0, weight_decay=0 only)
RMSPROP_EPS = 1e-8


def is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


def joints_mse_torch(outputs: List[torch.Tensor], target: torch.Tensor, target_weight: torch.Tensor,
                     use_target_weight: bool = True) -> torch.Tensor:
    """mse.py:27-44 in closed form: sum_s 1/(2*J*B*hw) * sum (w*(p-g))^2."""
    B, J = target.shape[:2]
    hw = target.shape[2] * target.shape[3]
    w = target_weight.reshape(B, J, 1, 1) if use_target_weight else torch.ones(B, J, 1, 1)
    loss = 0.0
    for o in outputs:
        d = (o - target) * w
        loss = loss + 0.5 * (d * d).sum() / (B * hw * J)
    return loss


def forward_backward(sd: Dict[str, torch.Tensor], x: torch.Tensor, target: torch.Tensor, target_weight: torch.Tensor,
                     use_target_weight: bool = True, grad_scale: float = 1.0
                     ) -> Tuple[float, List[torch.Tensor], Dict[str, torch.Tensor]]:
    """One train-mode forward + backward (trainer.py:89-98).  Updates the BN running statistics in `sd`
    in place, exactly as the reference's modules do.  Returns (loss, outputs, {param key: grad})."""
    leaves = {}
    work = {}
    for k, v in sd.items():
        if is_param(k):
            leaves[k] = v.detach().clone().requires_grad_(True)
            work[k] = leaves[k]
        else:
            work[k] = v                      # running stats: updated in place by F.batch_norm
    H._TRAINING[0] = True
    try:
        outs = H.hg_forward(work, x)
    finally:
        H._TRAINING[0] = False
    loss = joints_mse_torch(outs, target, target_weight, use_target_weight)
    (loss * grad_scale).backward()
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k] = work[k]
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in leaves.items()}
    return float(loss.detach()), [o.detach() for o in outs], grads


def rmsprop_update(sd: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], square_avg: Dict[str, torch.Tensor],
                   lr: float, alpha: float = RMSPROP_ALPHA, eps: float = RMSPROP_EPS) -> None:
    """torch.optim.RMSprop.step with momentum=0, centered=False, weight_decay=0:
         v <- alpha*v + (1-alpha)*g^2 ;  p <- p - lr * g / (sqrt(v) + eps).   In place on sd / square_avg."""
    for k, g in grads.items():
        v = square_avg.setdefault(k, torch.zeros_like(g))
        v.mul_(alpha).addcmul_(g, g, value=1 - alpha)
        sd[k] = sd[k] - lr * g / (v.sqrt() + eps)


def adjust_learning_rate(lr: float, epoch: int, schedule, gamma: float) -> float:
    """trainer.py:15-21: lr *= gamma when epoch is in the schedule."""
    return lr * gamma if epoch in schedule else lr


def train_steps(sd, batches, lr: float, use_target_weight: bool = True):
    """Runs len(batches) steps; returns per-step losses and the gradients of the last step."""
    state = {}
    losses, grads = [], None
    for x, target, tw in batches:
        loss, _, grads = forward_backward(sd, x, target, tw, use_target_weight)
        rmsprop_update(sd, grads, state, lr)
        losses.append(loss)
    return losses, grads, state
