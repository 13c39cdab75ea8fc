"""Building blocks with the reference's constructor signatures and state_dict keys
(reference: src/models/modules.py:6-99).  They are parameter CONTAINERS: the arithmetic runs in
libhgb200's sm_100a kernels, driven by hgb200.engine from the owning HourglassNet; calling a block
on its own routes through the same kernels via a one-block plan.
"""
import torch
import torch.nn as nn

__all__ = ['HGBottleneck', 'Hourglass']


class HGBottleneck(nn.Module):
    expansion = 2

    def __init__(self, inplanes, planes, stride=1, downsample=None, mobile=False):
        super(HGBottleneck, self).__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=True)
        self.bn2 = nn.BatchNorm2d(planes)
        if mobile:
            self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride,
                                   padding=1, bias=True, groups=planes)
        else:
            self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride,
                                   padding=1, bias=True)
        self.bn3 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 2, kernel_size=1, bias=True)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        """x: fp32 NCHW.  Eval-mode forward of ONE bottleneck through the sm_100a kernels (bf16 storage, fp32 accumulate).
        Train mode is served through the owning HourglassNet (whose plan keeps what backward needs); a stand-alone
        block has no autograd graph to hand its activations to, so it raises instead of silently running eval math."""
        from hgb200 import ops
        if self.training:
            raise RuntimeError("HGBottleneck.forward: stand-alone blocks run in eval mode only; train through HourglassNet "
                               "(hgb200.train keeps the saved activations and the backward plan per network)")
        if self.stride != 1:
            raise ValueError("stride != 1 is never built by the reference's hourglass (src/models/hourglass.py:45-58)")
        bw = _block_weights({"b." + k: v for k, v in self.state_dict().items()}, "b", x.device)
        xh = ops.nchw_to_nhwc_bf16(x.contiguous())
        return ops.nhwc_bf16_to_nchw(_run_block(bw, xh))


def _host(t):
    from hgb200.engine import _host_copy
    return _host_copy(t)


def _to_dev(bw, dev):
    for k, v in vars(bw).items():
        if torch.is_tensor(v):
            setattr(bw, k, v.to(dev).contiguous())
    return bw


def _block_weights(sd, prefix, dev):
    """Folded weights of one bottleneck (host fold, hgb200/fold.py) on `dev`."""
    from hgb200.fold import BlockWeights
    return _to_dev(BlockWeights({k: _host(v) for k, v in sd.items()}, prefix), dev)


def _run_block(bw, xh, up_low=None):
    """One folded bottleneck on NHWC bf16 (K1 with the bn1 prologue, K2, K3 + residual [+ upsample-add])."""
    from hgb200 import ops
    a2 = ops.conv_nhwc(xh, bw.w1, bw.b1, ksize=1, cout=bw.planes, relu=True, in_scale=bw.s1, in_shift=bw.t1)
    if bw.depthwise:
        a3 = ops.dwconv3x3(a2, bw.w2, bw.b2, relu=True)
    else:
        a3 = ops.conv_nhwc(a2, bw.w2, bw.b2, ksize=3, cout=bw.planes, relu=True)
    if bw.downsample:
        return ops.conv_nhwc(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, x2=xh, up_low=up_low)
    return ops.conv_nhwc(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, residual=xh, up_low=up_low)


class Hourglass(nn.Module):
    def __init__(self, block, num_blocks, planes, depth, mobile, skip_mode='concat'):
        super(Hourglass, self).__init__()
        self.depth = depth
        self.block = block
        self.mobile = mobile
        self.hg = self._make_hour_glass(block, num_blocks, planes, depth)
        assert skip_mode in ['sum', 'concat']
        if skip_mode == 'concat':
            self.concat_conv = nn.Conv2d(in_channels=planes * block.expansion * 2,
                                         out_channels=planes * block.expansion,
                                         kernel_size=1, padding=0, groups=2)

    def _make_residual(self, block, num_blocks, planes):
        layers = []
        for i in range(0, num_blocks):
            layers.append(block(planes * block.expansion, planes, mobile=self.mobile))
        return nn.Sequential(*layers)

    def _make_hour_glass(self, block, num_blocks, planes, depth):
        hg = []
        for i in range(depth):
            res = []
            for j in range(3):
                res.append(self._make_residual(block, num_blocks, planes))
            if i == 0:
                res.append(self._make_residual(block, num_blocks, planes))
            hg.append(nn.ModuleList(res))
        return nn.ModuleList(hg)

    def forward(self, x):
        """x: fp32 NCHW [B, 2*planes, H, W], H and W multiples of 2**depth.  Eval-mode forward of ONE hourglass
        (modules.py:80-99) through the sm_100a kernels: the same launches the HourglassNet plan issues for a stack --
        max-pool, bottleneck chains, the upsample-add in the up1 chain's last epilogue (or the two-GEMM form of
        concat_conv) -- issued eagerly.  Train mode goes through the owning HourglassNet."""
        from hgb200 import ops
        from hgb200.fold import chain_weights, concat_weights
        if self.training:
            raise RuntimeError("Hourglass.forward: stand-alone modules run in eval mode only; train through HourglassNet")
        if x.dim() != 4 or x.shape[2] % (1 << self.depth) or x.shape[3] % (1 << self.depth):
            raise ValueError(f"Hourglass.forward: expected [B,C,H,W] with H, W multiples of {1 << self.depth}")
        dev = x.device
        sd = {k: _host(v) for k, v in self.state_dict().items()}
        levels = [[[_to_dev(b, dev) for b in chain_weights(sd, f"hg.{d}.{k}")] for k in range(4 if d == 0 else 3)]
                  for d in range(self.depth)]
        cat = None
        if "concat_conv.weight" in sd:
            cat = [t.to(dev) for t in concat_weights(sd["concat_conv.weight"].float(), sd["concat_conv.bias"].float())]

        def chain(blocks, t, up_low=None):
            for i, bw in enumerate(blocks):
                t = _run_block(bw, t, up_low if i == len(blocks) - 1 else None)
            return t

        def level(d, t):
            low1 = chain(levels[d][1], ops.maxpool2x2(t))
            low2 = level(d - 1, low1) if d > 0 else chain(levels[0][3], low1)
            low3 = chain(levels[d][2], low2)
            if cat is None:
                return chain(levels[d][0], t, up_low=low3)
            wa, ba, wb, bb = cat
            up1 = chain(levels[d][0], t)
            low = ops.conv_nhwc(low3, wb, bb, ksize=1, cout=wb.shape[0])
            return ops.conv_nhwc(up1, wa, ba, ksize=1, cout=wa.shape[0], up_low=low)

        return ops.nhwc_bf16_to_nchw(level(self.depth - 1, ops.nchw_to_nhwc_bf16(x.contiguous())))
