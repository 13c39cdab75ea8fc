"""Weight preparation for the sm_100a path: eval-mode BatchNorm folding and GEMM weight layout.

The reference's bottleneck is PRE-activation (src/models/modules.py:27-47):
    conv1(relu(bn1(x))) -> conv2(relu(bn2(.))) -> conv3(relu(bn3(.))) + residual
so bn2 / bn3 FOLLOW conv1 / conv2 and fold into them (fp32, rounded to bf16 once), while bn1 sits
behind a ReLU on the raw residual stream and becomes the conv1 kernel's A-operand prologue
(per-channel scale/shift + ReLU applied to the TMA-landed tile).

GEMM weight layout expected by hg_conv_nhwc_bf16: bf16 [cout_pad][taps*cin (+cin2)], k = tap*cin + c.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

BN_EPS = 1e-5


def bn_scale_shift(sd: Dict[str, torch.Tensor], p: str) -> Tuple[torch.Tensor, torch.Tensor]:
    s = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + BN_EPS)
    t = sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * s
    return s.contiguous(), t.contiguous()


def fold_bn_after_conv(sd, conv_p: str, bn_p: Optional[str]):
    w = sd[conv_p + ".weight"].float()
    b = sd[conv_p + ".bias"].float()
    if bn_p is not None:
        s, t = bn_scale_shift(sd, bn_p)
        w = w * s.view(-1, 1, 1, 1)
        b = b * s + t
    return w, b


def gemm_weight(w4: torch.Tensor, extra: Optional[torch.Tensor] = None, k_pad: int = 0):
    """[cout, cin, kh, kw] (fp32) -> bf16 [cout_pad, kh*kw*cin (+extra cols)] K-major, rows padded to 16."""
    cout = w4.shape[0]
    m = w4.permute(0, 2, 3, 1).reshape(cout, -1)
    if extra is not None:
        m = torch.cat([m, extra], dim=1)
    if k_pad > m.shape[1]:
        m = torch.cat([m, m.new_zeros(cout, k_pad - m.shape[1])], dim=1)
    cout_pad = (cout + 15) // 16 * 16
    out = m.new_zeros(cout_pad, m.shape[1])
    out[:cout] = m
    return out.to(torch.bfloat16).contiguous()


def stem_window_weight(w4: torch.Tensor):
    """[64, 3, 7, 7] (fp32, BN already folded) -> bf16 [64, 224] for hg_stem_conv: k = ky*32 + (1+kx)*4 + c.
    The A operand of filter row ky is the 8-pixel x 4-channel window starting one pixel left of the
    first tap, so slots (pixel 0) and (channel 3) carry zero weights."""
    cout = w4.shape[0]
    m = w4.new_zeros(cout, 7, 8, 4)
    m[:, :, 1:8, 0:3] = w4.permute(0, 2, 3, 1)          # [cout, ky, kx, c]
    return m.reshape(cout, 224).to(torch.bfloat16).contiguous()


def pad_bias(b: torch.Tensor):
    cout = b.shape[0]
    cout_pad = (cout + 15) // 16 * 16
    out = b.new_zeros(cout_pad)
    out[:cout] = b
    return out.float().contiguous()


class BlockWeights:
    """One HGBottleneck, folded."""

    def __init__(self, sd, p: str):
        w2 = sd[p + ".conv2.weight"]
        # mobile=True: conv2 is depthwise (groups=planes, weight [planes,1,3,3]; src/models/modules.py:15-17)
        self.depthwise = w2.shape[1] == 1 and w2.shape[0] > 1
        self.cin = sd[p + ".conv1.weight"].shape[1]
        self.planes = sd[p + ".conv1.weight"].shape[0]
        self.cout = sd[p + ".conv3.weight"].shape[0]
        self.s1, self.t1 = bn_scale_shift(sd, p + ".bn1")
        w, b = fold_bn_after_conv(sd, p + ".conv1", p + ".bn2")
        self.w1, self.b1 = gemm_weight(w), pad_bias(b)
        w, b = fold_bn_after_conv(sd, p + ".conv2", p + ".bn3")
        if self.depthwise:
            self.w2, self.b2 = w.reshape(-1).contiguous(), b.contiguous()      # fp32 [c][9] for the CUDA-core stencil
        else:
            self.w2, self.b2 = gemm_weight(w), pad_bias(b)
        w, b = fold_bn_after_conv(sd, p + ".conv3", None)
        self.downsample = (p + ".downsample.0.weight") in sd
        if self.downsample:
            wd = sd[p + ".downsample.0.weight"].float()[:, :, 0, 0]
            b = b + sd[p + ".downsample.0.bias"].float()
            self.w3 = gemm_weight(w, extra=wd)       # [W3 | Wd]: second K segment reads the raw block input
        else:
            self.w3 = gemm_weight(w)
        self.b3 = pad_bias(b)


def concat_weights(w4: torch.Tensor, b: torch.Tensor):
    """skip_mode='concat' (src/models/modules.py:58-61,91-93): conv1x1(cat([up1, up2]), groups=2) with weight
    [2p, 2p, 1, 1] -- output channels [0,p) read up1, [p,2p) read up2 = upsample(low3).  A 1x1 convolution commutes
    with nearest upsampling, so the layer is two zero-padded 2p->2p GEMMs on the existing kernels:
        T   = [0 ; W_b] low3 + [0 ; b_b]            (at low resolution)
        out = [W_a ; 0] up1 + [b_a ; 0] + upsample(T)   (the upsample-add epilogue)
    Returns (Wa_pad, ba_pad, Wb_pad, bb_pad) in GEMM layout."""
    co, ci = w4.shape[0], w4.shape[1]
    half = co // 2
    w = w4[:, :, 0, 0]
    wa, wb = torch.zeros_like(w), torch.zeros_like(w)
    wa[:half], wb[half:] = w[:half], w[half:]
    ba, bb = torch.zeros_like(b), torch.zeros_like(b)
    ba[:half], bb[half:] = b[:half], b[half:]
    return (gemm_weight(wa[:, :, None, None]), pad_bias(ba), gemm_weight(wb[:, :, None, None]), pad_bias(bb))


def chain_weights(sd, p: str):
    blocks = []
    i = 0
    while f"{p}.{i}.bn1.weight" in sd:
        blocks.append(BlockWeights(sd, f"{p}.{i}"))
        i += 1
    if not blocks:
        raise KeyError(f"no bottleneck under '{p}'")
    return blocks


class NetWeights:
    """All folded weights of a HourglassNet state_dict (reference key layout, src/models/hourglass.py:9-43)."""

    def __init__(self, sd: Dict[str, torch.Tensor], depth: int = 4):
        sd = {k: v.detach() for k, v in sd.items()}
        self.depth = depth
        self.num_stacks = 0
        while f"score.{self.num_stacks}.weight" in sd:
            self.num_stacks += 1
        self.num_classes = sd["score.0.weight"].shape[0]
        w, b = fold_bn_after_conv(sd, "conv1", "bn1")
        self.stem_w, self.stem_b = gemm_weight(w, k_pad=192), pad_bias(b)     # im2col form (fallback / tests)
        self.stem_w_win = stem_window_weight(w)                                # window form (hg_stem_conv)
        self.layer1 = chain_weights(sd, "layer1")
        self.layer2 = chain_weights(sd, "layer2")
        self.layer3 = chain_weights(sd, "layer3")
        self.hg, self.res, self.fc, self.score, self.remap = [], [], [], [], []
        self.concat = []
        for i in range(self.num_stacks):
            if f"hg.{i}.concat_conv.weight" in sd:
                self.concat.append(concat_weights(sd[f"hg.{i}.concat_conv.weight"].float(),
                                                  sd[f"hg.{i}.concat_conv.bias"].float()))
            levels = []
            for d in range(depth):
                levels.append([chain_weights(sd, f"hg.{i}.hg.{d}.{k}") for k in range(4 if d == 0 else 3)])
            self.hg.append(levels)
            self.res.append(chain_weights(sd, f"res.{i}"))
            w, b = fold_bn_after_conv(sd, f"fc.{i}.0", f"fc.{i}.1")
            self.fc.append((gemm_weight(w), pad_bias(b)))
            ws = sd[f"score.{i}.weight"].float()
            bs = sd[f"score.{i}.bias"].float()
            self.score.append((gemm_weight(ws), pad_bias(bs)))
            if i < self.num_stacks - 1:
                # x + fc_(y) + score_(score(y)) = x + (W_fc_ + W_s_ W_s) y + (b_fc_ + W_s_ b_s + b_s_)
                # (src/models/hourglass.py:86-89): both remap convs are linear in y, so they merge into ONE
                # 256->256 GEMM whose epilogue adds the residual x.
                ws_ = sd[f"score_.{i}.weight"].float()[:, :, 0, 0]
                wm = sd[f"fc_.{i}.weight"].float()[:, :, 0, 0] + ws_ @ ws[:, :, 0, 0]
                bm = sd[f"fc_.{i}.bias"].float() + ws_ @ bs + sd[f"score_.{i}.bias"].float()
                self.remap.append((gemm_weight(wm[:, :, None, None]), pad_bias(bm)))
