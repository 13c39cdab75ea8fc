"""Shared GPU test helpers: torch fp32 references of each fused op and the conv case table.

Used by tests/test_gpu_*.py (pytest, -m gpu) and tools/gpu_debug.py (a verbose sweep that keeps going
after a failure so one gpurun call yields a full picture).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


# name, n, h, w, cin, cout, ksize, relu, prologue, residual, up_low, cin2, heads
CONV_CASES = [
    ("1x1_single_tile", 1, 8, 16, 64, 64, 1, False, False, False, False, 0, False),
    ("1x1_k256_n128_relu", 2, 16, 16, 256, 128, 1, True, False, False, False, 0, False),
    ("1x1_ragged_m", 3, 10, 6, 128, 128, 1, False, False, False, False, 0, False),
    ("1x1_prologue", 3, 10, 6, 256, 128, 1, True, True, False, False, 0, False),
    ("1x1_prologue_n64", 2, 16, 16, 64, 64, 1, True, True, False, False, 0, False),
    ("1x1_residual_n256", 2, 16, 16, 128, 256, 1, False, False, True, False, 0, False),
    ("1x1_residual_up", 2, 16, 16, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_two_inputs", 2, 16, 16, 64, 128, 1, False, False, False, False, 64, False),
    ("1x1_n256_k256_res", 2, 32, 32, 256, 256, 1, False, False, True, False, 0, False),
    ("3x3_64x64", 2, 64, 64, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_32x32", 3, 32, 32, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_16x16", 3, 16, 16, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_8x8", 5, 8, 8, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_4x4", 21, 4, 4, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_64x48", 2, 64, 48, 128, 128, 3, True, False, False, False, 0, False),
    ("3x3_6x12", 3, 6, 12, 64, 64, 3, False, False, False, False, 0, False),
    ("3x3_128x128_c64", 1, 128, 128, 64, 64, 3, True, False, False, False, 0, False),
    ("3x3_3x3img", 7, 3, 3, 128, 128, 3, True, False, False, False, 0, False),
    ("heads_j16", 2, 64, 64, 256, 16, 1, False, False, False, False, 0, True),
    ("heads_j17", 2, 64, 48, 256, 17, 1, False, False, False, False, 0, True),
    ("heads_j21", 1, 16, 16, 256, 21, 1, False, False, False, False, 0, True),
    ("heads_j14", 3, 8, 8, 256, 14, 1, False, False, False, False, 0, True),
    ("1x1_up_32x32", 3, 32, 32, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_8x8", 6, 8, 8, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_8x8_odd_n", 5, 8, 8, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_4x4", 9, 4, 4, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_64x48_generic", 2, 64, 48, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_6x6_generic", 3, 6, 6, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_22x24_rows", 3, 22, 24, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_64x96_wide", 1, 64, 96, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_up_10x12_relu_rows", 5, 10, 12, 256, 256, 1, True, False, True, True, 0, False),
    ("1x1_up_2x2", 5, 2, 2, 128, 256, 1, False, False, True, True, 0, False),
    ("1x1_ragged_res_n256", 3, 10, 6, 128, 256, 1, False, False, True, False, 0, False),
    ("1x1_two_inputs_n256", 2, 32, 32, 128, 256, 1, False, False, False, False, 128, False),
    ("1x1_remap_relu", 2, 64, 64, 256, 256, 1, True, False, True, False, 0, False),
    ("1x1_k192_n64_stem", 1, 128, 128, 192, 64, 1, True, False, False, False, 0, False),
    ("1x1_many_tiles", 8, 64, 64, 256, 128, 1, True, True, False, False, 0, False),
    ("3x3_many_tiles", 6, 64, 64, 128, 128, 3, True, False, False, False, 0, False),
    ("1x1_n256_many_tiles", 6, 64, 64, 128, 256, 1, False, False, True, True, 0, False),
]


def make_conv_case(case, device, seed=0):
    (name, n, h, w, cin, cout, ksize, relu, prologue, residual, up_low, cin2, heads) = case
    g = torch.Generator(device="cpu").manual_seed(seed + sum(map(ord, name)) % 1000)
    taps = ksize * ksize
    cout_pad = (cout + 15) // 16 * 16
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, ksize, ksize, generator=g) / math.sqrt(cin * taps)
    bias = torch.randn(cout, generator=g) * 0.5
    t = dict(name=name, n=n, h=h, w=w, cin=cin, cout=cout, ksize=ksize, relu=relu, heads=heads)
    t["x"] = x.to(torch.bfloat16).to(device)
    w_mat = wt.permute(0, 2, 3, 1).reshape(cout, taps * cin)
    if cin2:
        x2 = torch.randn(n, h, w, cin2, generator=g)
        w2 = torch.randn(cout, cin2, generator=g) / math.sqrt(cin2)
        t["x2"] = x2.to(torch.bfloat16).to(device)
        w_mat = torch.cat([w_mat, w2], dim=1)
        t["w2"] = w2
    w_pad = torch.zeros(cout_pad, w_mat.shape[1])
    w_pad[:cout] = w_mat
    b_pad = torch.zeros(cout_pad)
    b_pad[:cout] = bias
    t["weight"] = w_pad.to(torch.bfloat16).to(device)
    t["bias"] = b_pad.to(device)
    t["wt4"] = wt
    if prologue:
        t["in_scale"] = (0.5 + torch.rand(cin, generator=g)).to(device)
        t["in_shift"] = (0.3 * torch.randn(cin, generator=g)).to(device)
    if residual:
        t["residual"] = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16).to(device)
    if up_low:
        t["up_low"] = torch.randn(n, h // 2, w // 2, cout, generator=g).to(torch.bfloat16).to(device)
    return t


def conv_reference(t):
    """fp32 torch reference of the fused op, from the SAME bf16-rounded operands."""
    no_tf32()
    dev = t["x"].device
    x = t["x"].float()
    if "in_scale" in t:
        x = r16(F.relu(x * t["in_scale"] + t["in_shift"]))       # prologue result is a bf16 MMA operand
    cout, ksize = t["cout"], t["ksize"]
    taps = ksize * ksize
    cin = t["cin"]
    wmat = t["weight"].float()[:cout]
    w4 = wmat[:, :taps * cin].reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(x.permute(0, 3, 1, 2), w4.to(dev), t["bias"][:cout], padding=ksize // 2)
    if "x2" in t:
        w2 = wmat[:, taps * cin:]
        y = y + F.conv2d(t["x2"].float().permute(0, 3, 1, 2), w2[:, :, None, None].to(dev))
    if "residual" in t:
        y = y + t["residual"].float().permute(0, 3, 1, 2)
    if "up_low" in t:
        y = y + F.interpolate(t["up_low"].float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    if t["relu"]:
        y = F.relu(y)
    return y          # NCHW fp32


def run_conv_case(t):
    from hgb200 import ops
    out = ops.conv_nhwc(t["x"], t["weight"], t["bias"], ksize=t["ksize"], cout=t["cout"], relu=t["relu"],
                        in_scale=t.get("in_scale"), in_shift=t.get("in_shift"), residual=t.get("residual"),
                        up_low=t.get("up_low"), x2=t.get("x2"), heads=t["heads"])
    torch.cuda.synchronize()
    ops.check_err_word(t["x"].device)
    if t["heads"]:
        return out
    return out.float().permute(0, 3, 1, 2)


def conv_error(out, ref):
    """max |out - ref| relative to max |ref|, plus the bf16 output-rounding allowance."""
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    return err, scale
