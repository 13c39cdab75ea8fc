"""fp32 CPU restatement of the reference stacked-hourglass forward.

Test infrastructure only (see oracle/__init__.py).  The network is expressed as
pure functions over a flat ``state_dict`` carrying the reference's exact key
names (checkpoint row), so the same weights drive the
oracle, the live reference and the B200 path.

Reference followed:
  * HourglassNet.forward          src/models/hourglass.py:69-90
  * HourglassNet.__init__ layout  src/models/hourglass.py:9-43
  * HGBottleneck.forward          src/models/modules.py:27-47
  * Hourglass._hour_glass_forward src/models/modules.py:80-96
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default used everywhere in the reference


_CALIBRATE = [False]
_TRAINING = [False]      # train-mode BatchNorm (batch statistics + running-stat update), see oracle/train_oracle.py
BN_MOMENTUM = 0.1        # nn.BatchNorm2d default used everywhere in the reference


def _bn(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """BatchNorm2d: running statistics in eval mode; batch statistics (biased variance for the
    normalisation, unbiased for the running-variance update, momentum 0.1) in train mode."""
    if _TRAINING[0]:
        sd[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                            sd[p + ".weight"], sd[p + ".bias"], True, BN_MOMENTUM, BN_EPS)
    if _CALIBRATE[0]:   # see calibrate_bn(): overwrite running stats with this batch's statistics
        sd[p + ".running_mean"] = x.mean(dim=(0, 2, 3))
        sd[p + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv(sd, p, x, padding=0, stride=1, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride,
                    padding=padding, groups=groups)


def bottleneck(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """Pre-activation bottleneck, src/models/modules.py:27-47.

    ``mobile`` (depthwise conv2, modules.py:15-17) is recognised from the
    weight shape: a depthwise 3x3 has weight [planes, 1, 3, 3].
    """
    out = F.relu(_bn(sd, p + ".bn1", x))
    out = _conv(sd, p + ".conv1", out)
    out = F.relu(_bn(sd, p + ".bn2", out))
    w2 = sd[p + ".conv2.weight"]
    groups = w2.shape[0] if w2.shape[1] == 1 and w2.shape[0] > 1 else 1
    out = _conv(sd, p + ".conv2", out, padding=1, groups=groups)
    out = F.relu(_bn(sd, p + ".bn3", out))
    out = _conv(sd, p + ".conv3", out)
    if (p + ".downsample.0.weight") in sd:          # hourglass.py:46-51
        residual = _conv(sd, p + ".downsample.0", x)
    else:
        residual = x
    return out + residual


def _residual_chain(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential of bottlenecks `p.0`, `p.1`, ... (num_blocks of them)."""
    i = 0
    while (f"{p}.{i}.bn1.weight") in sd:
        x = bottleneck(sd, f"{p}.{i}", x)
        i += 1
    assert i > 0, f"no bottleneck under {p}"
    return x


def hourglass(sd, p: str, n: int, x: torch.Tensor) -> torch.Tensor:
    """Recursive encoder/decoder, src/models/modules.py:80-96."""
    up1 = _residual_chain(sd, f"{p}.hg.{n-1}.0", x)
    low1 = F.max_pool2d(x, 2, stride=2)
    low1 = _residual_chain(sd, f"{p}.hg.{n-1}.1", low1)
    if n > 1:
        low2 = hourglass(sd, p, n - 1, low1)
    else:
        low2 = _residual_chain(sd, f"{p}.hg.{n-1}.3", low1)
    low3 = _residual_chain(sd, f"{p}.hg.{n-1}.2", low2)
    up2 = F.interpolate(low3, scale_factor=2, mode="nearest")
    if (p + ".concat_conv.weight") in sd:           # skip_mode == 'concat'
        out = torch.cat([up1, up2], dim=1)
        return _conv(sd, p + ".concat_conv", out, groups=2)
    return up1 + up2


def num_stacks_of(sd) -> int:
    s = 0
    while f"score.{s}.weight" in sd:
        s += 1
    return s


def hg_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor,
               depth: int = 4) -> List[torch.Tensor]:
    """HourglassNet.forward, src/models/hourglass.py:69-90 (eval mode)."""
    S = num_stacks_of(sd)
    out = []
    x = _conv(sd, "conv1", x, padding=3, stride=2)
    x = F.relu(_bn(sd, "bn1", x))
    x = _residual_chain(sd, "layer1", x)
    x = F.max_pool2d(x, 2, stride=2)
    x = _residual_chain(sd, "layer2", x)
    x = _residual_chain(sd, "layer3", x)
    for i in range(S):
        y = hourglass(sd, f"hg.{i}", depth, x)
        y = _residual_chain(sd, f"res.{i}", y)
        y = F.relu(_bn(sd, f"fc.{i}.1", _conv(sd, f"fc.{i}.0", y)))
        score = _conv(sd, f"score.{i}", y)
        out.append(score)
        if i < S - 1:
            x = x + _conv(sd, f"fc_.{i}", y) + _conv(sd, f"score_.{i}", score)
    return out


# ---------------------------------------------------------------------------
# Deterministic synthetic weights with the reference's state_dict layout.
# ---------------------------------------------------------------------------

def _conv_init(g, cout, cin_per_group, k):
    """nn.Conv2d's default init (kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    for both weight and bias), seeded -- SURVEY.md section 8d "default init"."""
    bound = 1.0 / math.sqrt(cin_per_group * k * k)
    w = (torch.rand(cout, cin_per_group, k, k, generator=g) * 2 - 1) * bound
    b = (torch.rand(cout, generator=g) * 2 - 1) * bound
    return w, b


def make_state_dict(num_stacks=2, num_blocks=1, num_classes=16, mobile=False,
                    skip_mode="sum", seed=0, randomize_bn=True
                    ) -> Dict[str, torch.Tensor]:
    """A seeded state_dict with exactly the reference's keys and shapes.

    Key layout follows HourglassNet.__init__ (hourglass.py:9-43) and
    Hourglass._make_hour_glass (modules.py:69-78).  BN running statistics and
    affine parameters are randomised so that folding
    them into the convolutions is a non-trivial identity.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(p, cin, cout, k, groups=1):
        w, b = _conv_init(g, cout, cin // groups, k)
        sd[p + ".weight"], sd[p + ".bias"] = w, b

    def bn(p, c):
        if randomize_bn:
            sd[p + ".weight"] = 0.75 + 0.5 * torch.rand(c, generator=g)
            sd[p + ".bias"] = 0.1 * torch.randn(c, generator=g)
            sd[p + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[p + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[p + ".weight"] = torch.ones(c)
            sd[p + ".bias"] = torch.zeros(c)
            sd[p + ".running_mean"] = torch.zeros(c)
            sd[p + ".running_var"] = torch.ones(c)
        sd[p + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def block(p, inplanes, planes, downsample=False, mob=False):
        bn(p + ".bn1", inplanes)
        conv(p + ".conv1", inplanes, planes, 1)
        bn(p + ".bn2", planes)
        conv(p + ".conv2", planes, planes, 3, groups=planes if mob else 1)
        bn(p + ".bn3", planes)
        conv(p + ".conv3", planes, planes * 2, 1)
        if downsample:
            conv(p + ".downsample.0", inplanes, planes * 2, 1)

    def residual(p, inplanes, planes, blocks, first_mobile):
        # hourglass.py:45-58: only the first block gets `mobile`/downsample
        block(f"{p}.0", inplanes, planes, downsample=(inplanes != planes * 2), mob=first_mobile)
        for i in range(1, blocks):
            block(f"{p}.{i}", planes * 2, planes)

    conv("conv1", 3, 64, 7)
    bn("bn1", 64)
    # hourglass.py:21-23: planes = (inplanes=64, inplanes=128, num_feats=128)
    residual("layer1", 64, 64, 1, mobile)      # 64 -> 128, downsample
    residual("layer2", 128, 128, 1, mobile)    # 128 -> 256, downsample (at 64x64)
    residual("layer3", 256, 128, 1, mobile)    # 256 -> 256, identity residual
    ch = 256
    for i in range(num_stacks):
        for d in range(4):
            for k in range(4 if d == 0 else 3):
                for b in range(num_blocks):   # modules.py:63-67: every block gets mobile
                    block(f"hg.{i}.hg.{d}.{k}.{b}", ch, 128, mob=mobile)
        if skip_mode == "concat":
            conv(f"hg.{i}.concat_conv", 2 * ch, ch, 1, groups=2)
        residual(f"res.{i}", ch, 128, num_blocks, mobile)
        conv(f"fc.{i}.0", ch, ch, 1)
        bn(f"fc.{i}.1", ch)
        conv(f"score.{i}", ch, num_classes, 1)
        if i < num_stacks - 1:
            conv(f"fc_.{i}", ch, ch, 1)
            conv(f"score_.{i}", num_classes, ch, 1)
    return sd


def calibrate_bn(sd, x: torch.Tensor):
    """Replace every BN's running statistics by the batch statistics seen on `x`.

    Random conv weights with arbitrary running stats make activations grow ~10x per
    stack; calibrating gives a well-conditioned network (unit-variance pre-activations,
    like a trained one) whose weights are still fully synthetic and seeded.  In place.
    """
    _CALIBRATE[0] = True
    try:
        with torch.no_grad():
            hg_forward(sd, x)
    finally:
        _CALIBRATE[0] = False
    return sd


def conv_flops_per_image(sd, h=256, w=256) -> float:
    """2*MAC over all convs for one forward (SURVEY.md section 6 probe)."""
    x = torch.zeros(1, 3, h, w)
    total = [0.0]
    orig = F.conv2d

    def counting(inp, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        out = orig(inp, weight, bias, stride, padding, dilation, groups)
        cout, cin_g, kh, kw = weight.shape
        total[0] += 2.0 * out.shape[0] * out.shape[2] * out.shape[3] * cout * cin_g * kh * kw
        return out

    F.conv2d = counting
    try:
        with torch.no_grad():
            hg_forward(sd, x)
    finally:
        F.conv2d = orig
    return total[0]
