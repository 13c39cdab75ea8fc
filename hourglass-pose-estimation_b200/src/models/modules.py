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
        """x: fp32 NCHW.  Eval-mode forward of ONE bottleneck through the sm_100a kernels
        (bf16 storage, fp32 accumulate)."""
        from hgb200 import ops
        from hgb200.fold import BlockWeights
        if self.training:
            raise NotImplementedError("stand-alone HGBottleneck.forward is eval-only; train through HourglassNet")
        if self.stride != 1:
            raise NotImplementedError("stride != 1 is never used by the reference's hourglass")
        sd = {"b." + k: v.detach() for k, v in self.state_dict().items()}
        bw = BlockWeights(sd, "b")
        dev = x.device
        for k, v in vars(bw).items():
            if torch.is_tensor(v):
                setattr(bw, k, v.to(dev).contiguous())
        xh = ops.nchw_to_nhwc_bf16(x.contiguous())
        a2 = ops.conv_nhwc(xh, bw.w1, bw.b1, ksize=1, cout=bw.planes, relu=True, in_scale=bw.s1, in_shift=bw.t1)
        a3 = ops.conv_nhwc(a2, bw.w2, bw.b2, ksize=3, cout=bw.planes, relu=True)
        if bw.downsample:
            out = ops.conv_nhwc(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, x2=xh)
        else:
            out = ops.conv_nhwc(a3, bw.w3, bw.b3, ksize=1, cout=bw.cout, residual=xh)
        return ops.nhwc_bf16_to_nchw(out)


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
        raise NotImplementedError("Hourglass is executed as part of HourglassNet's fused plan (hgb200.engine)")
