"""CPU emulation of the B200 numeric path (bf16 storage, fp32 accumulate).

Test infrastructure only.  Mirrors, in plain torch on the CPU, WHERE the CUDA
path rounds to bf16: folded weights are rounded once, every tensor written to
HBM between kernels is bf16, accumulation / bias / BN-prologue / residual adds
are fp32 inside a kernel.  It predicts the error budget of the design against
the fp32 oracle and gives a per-layer checker for the kernels
that is tight to ~1 bf16 ulp, which the fp32 oracle cannot.

Math follows the reference exactly (same citations as hourglass_oracle.py);
only the rounding points are the B200 design's.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from .hourglass_oracle import BN_EPS, num_stacks_of


def r16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def rw(w: torch.Tensor) -> torch.Tensor:
    """Rounding applied to (folded) weights -- separate hook so experiments can toggle it."""
    return r16(w)


def bn_scale_shift(sd, p):
    s = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    t = sd[p + ".bias"] - sd[p + ".running_mean"] * s
    return s, t


def fold_following_bn(sd, conv_p, bn_p):
    """conv followed by BN (eval): fold BN into the conv's weight/bias (fp32)."""
    s, t = bn_scale_shift(sd, bn_p)
    w = sd[conv_p + ".weight"] * s.view(-1, 1, 1, 1)
    b = sd[conv_p + ".bias"] * s + t
    return w, b


class Emu:
    def __init__(self, sd: Dict[str, torch.Tensor], residual_fp32: bool = False):
        self.sd = sd
        self.rs = (lambda x: x) if residual_fp32 else r16   # rounding of the raw residual stream

    def block(self, p, x):
        sd = self.sd
        s1, t1 = bn_scale_shift(sd, p + ".bn1")
        a1 = r16(F.relu(x * s1.view(1, -1, 1, 1) + t1.view(1, -1, 1, 1)))     # K1 prologue, bf16 MMA operand
        w1, b1 = fold_following_bn(sd, p + ".conv1", p + ".bn2")
        a2 = r16(F.relu(F.conv2d(a1, rw(w1), b1)))
        w2, b2 = fold_following_bn(sd, p + ".conv2", p + ".bn3")
        groups = w2.shape[0] if w2.shape[1] == 1 and w2.shape[0] > 1 else 1
        a3 = r16(F.relu(F.conv2d(a2, rw(w2), b2, padding=1, groups=groups)))
        out = F.conv2d(a3, rw(sd[p + ".conv3.weight"]), sd[p + ".conv3.bias"])
        if (p + ".downsample.0.weight") in sd:
            out = out + F.conv2d(r16(x), rw(sd[p + ".downsample.0.weight"]), sd[p + ".downsample.0.bias"])
        else:
            out = out + x
        return out            # caller decides rounding (fused epilogues add more terms first)

    def chain(self, p, x):
        i = 0
        while f"{p}.{i}.bn1.weight" in self.sd:
            x = self.rs(self.block(f"{p}.{i}", x))
            i += 1
        return x

    def chain_last_unrounded(self, p, x):
        """Like chain(), but the last block's sum is returned before rounding so a fused
        epilogue term (upsample-add) can be added in fp32 first."""
        n = 0
        while f"{p}.{n}.bn1.weight" in self.sd:
            n += 1
        for i in range(n - 1):
            x = self.rs(self.block(f"{p}.{i}", x))
        return self.block(f"{p}.{n-1}", x)

    def hourglass(self, p, n, x):
        sd = self.sd
        low1 = F.max_pool2d(x, 2, stride=2)
        low1 = self.chain(f"{p}.hg.{n-1}.1", low1)
        if n > 1:
            low2 = self.hourglass(p, n - 1, low1)
        else:
            low2 = self.chain(f"{p}.hg.{n-1}.3", low1)
        low3 = self.chain(f"{p}.hg.{n-1}.2", low2)
        up2 = F.interpolate(low3, scale_factor=2, mode="nearest")
        if (p + ".concat_conv.weight") in sd:
            up1 = self.chain(f"{p}.hg.{n-1}.0", x)
            out = torch.cat([r16(up1), r16(up2)], dim=1)
            return self.rs(F.conv2d(out, rw(sd[p + ".concat_conv.weight"]), sd[p + ".concat_conv.bias"], groups=2))
        # upsample-add fused into the up1 block's last conv3 epilogue (fp32 add, one rounding)
        return self.rs(self.chain_last_unrounded(f"{p}.hg.{n-1}.0", x) + up2)

    def forward(self, x, depth=4) -> List[torch.Tensor]:
        sd = self.sd
        S = num_stacks_of(sd)
        w, b = fold_following_bn(sd, "conv1", "bn1")
        x = r16(F.relu(F.conv2d(r16(x), rw(w), b, stride=2, padding=3)))
        x = self.chain("layer1", x)
        x = F.max_pool2d(x, 2, stride=2)
        x = self.chain("layer2", x)
        x = self.chain("layer3", x)
        outs = []
        for i in range(S):
            y = self.hourglass(f"hg.{i}", depth, x)
            y = self.chain(f"res.{i}", y)
            w, b = fold_following_bn(sd, f"fc.{i}.0", f"fc.{i}.1")
            y = r16(F.relu(F.conv2d(r16(y), rw(w), b)))
            score = F.conv2d(y, rw(sd[f"score.{i}.weight"]), sd[f"score.{i}.bias"])   # fp32 heat map out
            outs.append(score)
            if i < S - 1:
                # x + fc_(y) + score_(score(y)) == x + (W_fc_ + W_s_ W_s) y + (b_fc_ + W_s_ b_s + b_s_)
                ws_ = sd[f"score_.{i}.weight"][:, :, 0, 0]
                ws = sd[f"score.{i}.weight"][:, :, 0, 0]
                wm = sd[f"fc_.{i}.weight"][:, :, 0, 0] + ws_ @ ws
                bm = sd[f"fc_.{i}.bias"] + ws_ @ sd[f"score.{i}.bias"] + sd[f"score_.{i}.bias"]
                x = self.rs(x + F.conv2d(y, rw(wm)[:, :, None, None], bm))
        return outs


def emulate_forward(sd, x, residual_fp32=False):
    with torch.no_grad():
        return Emu(sd, residual_fp32).forward(x)
