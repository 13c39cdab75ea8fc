"""Tensor-level wrappers over the C ABI.  torch supplies memory and the current stream only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import lib, ConvDesc, PackEntry, HgError

_err_words = {}


def _require_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise HgError("hgb200 ops run on CUDA tensors only (no CPU fallback)")
        if not t.is_contiguous():
            raise HgError("hgb200 ops need contiguous tensors")


def require_device(device):
    """The path has no CPU fallback: engines call this before they allocate anything."""
    if torch.device(device).type != "cuda":
        raise HgError(f"hgb200 runs on CUDA devices only (got '{device}'): there is no CPU fallback; move the model with "
                      f".to('cuda')")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def err_word(device=None) -> torch.Tensor:
    """Per-device uint32 word that kernels set when a bounded mbarrier wait times out."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    w = _err_words.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=dev)
        _err_words[key] = w
    return w


def check_err_word(device=None):
    """Synchronises and raises if any kernel reported a protocol timeout."""
    w = err_word(device)
    v = int(w.item())
    if v != 0:
        w.zero_()
        raise HgError(f"device-side protocol timeout, code 0x{v & 0xffffffff:x} (role<<8 | barrier)")


def conv_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, *, ksize: int, cout: int,
              relu: bool = False, in_scale: Optional[torch.Tensor] = None, in_shift: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None, up_low: Optional[torch.Tensor] = None,
              x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              out_nchw_f32: Optional[torch.Tensor] = None, heads: bool = False,
              out_halo: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
              pool_out: Optional[torch.Tensor] = None, pool_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Implicit-GEMM conv on tcgen05 (hg_conv_nhwc_bf16).

    x: bf16 [n,h,w,cin]; weight: bf16 [cout_pad, taps*cin (+cin2)]; bias fp32 [cout_pad].
    heads=True -> returns fp32 NCHW [n,cout,h,w]; else bf16 NHWC [n,h,w,cout].
    stats: fp32 [2*cout], per-channel sum | sum of squares of the result, added by the kernel's epilogue (the
    batch statistics of the train-mode BatchNorm that follows).
    pool_out: bf16 [n,h/2,w/2,cout], additionally receives max_pool2d(result, 2, 2) from the same epilogue
    (see conv_pool_fusable).
    pool_in: bf16 [n,h/2,w/2,cin], additionally receives max_pool2d(x, 2, 2) of the RAW input from the prologue warps of a
    1x1 conv with in_scale / in_shift (see conv_pool_in_fusable).
    """
    _require_cuda(x, weight, bias, in_scale, in_shift, residual, up_low, x2, out, out_nchw_f32, stats)
    _check_stats(stats, cout)
    if stats is not None and heads:
        raise HgError("conv_nhwc: stats are for bf16 NHWC outputs")
    if pool_out is not None:
        _require_cuda(pool_out)
        if (not conv_pool_fusable(x.shape[1], x.shape[2], cout, weight.shape[1]) or ksize != 1 or heads or in_scale is not None
                or stats is not None or out_halo is not None):
            raise HgError("conv_nhwc: pool_out needs a plain 1x1 conv with cout 256 on a level conv_pool_fusable() accepts")
        if tuple(pool_out.shape) != (x.shape[0], x.shape[1] // 2, x.shape[2] // 2, cout) or pool_out.dtype != torch.bfloat16:
            raise HgError("conv_nhwc: pool_out must be bf16 [n,h/2,w/2,cout]")
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16 or (bias is not None and bias.dtype != torch.float32):
        raise HgError("conv_nhwc: x/weight must be bf16 and bias fp32")
    if pool_in is not None:
        _require_cuda(pool_in)
        if (not conv_pool_in_fusable(x.shape[0], x.shape[1], x.shape[2], cout) or ksize != 1 or heads or in_scale is None
                or stats is not None or pool_out is not None or x2 is not None):
            raise HgError("conv_nhwc: pool_in needs a 1x1 conv with prologue and cout 128 on a level conv_pool_in_fusable() accepts")
        if tuple(pool_in.shape) != (x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]) or pool_in.dtype != torch.bfloat16:
            raise HgError("conv_nhwc: pool_in must be bf16 [n,h/2,w/2,cin]")
    n, h, w, cin = x.shape
    cin2 = 0 if x2 is None else x2.shape[-1]
    cout_pad = (cout + 15) // 16 * 16
    ktot = ksize * ksize * cin + cin2
    if tuple(weight.shape) != (cout_pad, ktot) or (bias is not None and bias.numel() < cout_pad):
        raise HgError(f"conv_nhwc: weight {tuple(weight.shape)} / bias {None if bias is None else bias.numel()} do not "
                      f"match [{cout_pad},{ktot}]")
    d = ConvDesc()
    if out_halo is not None:
        _require_cuda(out_halo)
        if out_halo.dtype != torch.bfloat16 or out_halo.numel() != halo_padded_elems(n, h, w, cout) or heads:
            raise HgError("conv_nhwc: out_halo must be a bf16 buffer of halo_padded_elems(n,h,w,cout) elements")
        out = out_halo
        d.out_halo = 1
    if heads:
        if out_nchw_f32 is None:
            out_nchw_f32 = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
        result = out_nchw_f32
    else:
        if out is None:
            out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=x.device)
        result = out
    if ksize != 1 and out_halo is not None:
        raise HgError("conv_nhwc: out_halo is a 1x1-conv output mode")
    for t, shape in ((residual, (n, h, w, cout)), (up_low, (n, h // 2, w // 2, cout)), (x2, (n, h, w, cin2))):
        if t is not None and (tuple(t.shape) != shape or t.dtype != torch.bfloat16):
            raise HgError(f"conv_nhwc: operand shape {tuple(t.shape)} != {shape} or not bf16")
    d.in_, d.in2, d.weight = x.data_ptr(), (x2.data_ptr() if x2 is not None else None), weight.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.in_scale = in_scale.data_ptr() if in_scale is not None else None
    d.in_shift = in_shift.data_ptr() if in_shift is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.up_low = up_low.data_ptr() if up_low is not None else None
    d.out = None if heads else out.data_ptr()
    d.out_nchw_f32 = out_nchw_f32.data_ptr() if heads else None
    d.err_word = err_word(x.device).data_ptr()
    d.stats = stats.data_ptr() if stats is not None else None
    d.pool_out = pool_out.data_ptr() if pool_out is not None else None
    d.pool_in = pool_in.data_ptr() if pool_in is not None else None
    d.n, d.h, d.w, d.cin, d.cin2, d.cout, d.ksize, d.relu = n, h, w, cin, cin2, cout, ksize, int(relu)
    lib.check(lib.hg_conv_nhwc_bf16(C.byref(d), _stream()), "hg_conv_nhwc_bf16")
    return result


def conv_pool_fusable(h: int, w: int, cout: int, k: int) -> bool:
    """Whether hg_conv_nhwc_bf16 can also write the 2x2 max-pool of a 1x1 conv's [n,h,w,cout] result (pool_out): a
    128-pixel tile must hold whole pooling windows, and K = cin (+cin2) <= 128 so that the pooled staging slabs fit
    beside the resident weights without shrinking the staging ring."""
    return cout == 256 and k <= 128 and 2 <= w <= 64 and (w & (w - 1)) == 0 and h % 2 == 0 and 128 % (2 * w) == 0


def conv_pool_in_fusable(n: int, h: int, w: int, cout: int) -> bool:
    """Whether a 1x1 conv with prologue can also write the 2x2 max-pool of its RAW input (pool_in): flat 128-pixel tiles made
    of whole pooling windows, 128 output channels."""
    return (cout == 128 and 2 <= w <= 64 and (w & (w - 1)) == 0 and h % 2 == 0 and 128 % (2 * w) == 0
            and (n * h * w) % 128 == 0 and (h * w) % 128 == 0)


def stem_im2col(x_nchw: torch.Tensor, flip_w: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 NCHW [n,3,h,w] -> bf16 [n, h/2, w/2, 192] im2col rows of the 7x7/s2 stem conv."""
    _require_cuda(x_nchw, out)
    n, c, h, w = x_nchw.shape
    if c != 3 or x_nchw.dtype != torch.float32:
        raise HgError("stem_im2col: expects fp32 [n,3,h,w]")
    if out is None:
        out = torch.empty((n, h // 2, w // 2, 192), dtype=torch.bfloat16, device=x_nchw.device)
    lib.check(lib.hg_stem_im2col(_ptr(x_nchw), _ptr(out), n, h, w, int(flip_w), _stream()), "hg_stem_im2col")
    return out


def halo_padded_elems(n: int, h: int, w: int, c: int) -> int:
    return (n * (h + 1) * (w + 1) + (w + 1)) * c


def halo_padded_buffer(n: int, h: int, w: int, c: int, device) -> torch.Tensor:
    """Zero-initialised flat bf16 buffer in the halo-padded layout [zero row][n][h+1][w+1][c]."""
    return torch.zeros(halo_padded_elems(n, h, w, c), dtype=torch.bfloat16, device=device)


def halo_interior(buf: torch.Tensor, n: int, h: int, w: int, c: int) -> torch.Tensor:
    """Strided [n,h,w,c] view of the interior of a halo-padded buffer (tests / debugging)."""
    return buf[(w + 1) * c:].view(n, h + 1, w + 1, c)[:, :h, :w, :]


def _check_stats(stats: Optional[torch.Tensor], cout: int):
    if stats is not None and (stats.dtype != torch.float32 or stats.numel() < 2 * cout or not stats.is_cuda):
        raise HgError(f"stats must be a CUDA fp32 tensor of at least 2*cout = {2 * cout} elements")


def conv3x3_halo(x_halo: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, *, n: int, h: int, w: int, cin: int,
                 cout: int, relu: bool = False, out: Optional[torch.Tensor] = None,
                 stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 conv reading a halo-padded input (hg_conv3x3_halo_bf16) -> dense bf16 NHWC [n,h,w,cout].
    stats: fp32 [2*cout], per-channel sum | sum of squares of the result, added by the kernel's epilogue."""
    _require_cuda(x_halo, weight, bias, out, stats)
    _check_stats(stats, cout)
    if x_halo.numel() != halo_padded_elems(n, h, w, cin) or x_halo.dtype != torch.bfloat16:
        raise HgError("conv3x3_halo: input is not a halo-padded bf16 buffer of the stated shape")
    if tuple(weight.shape) != (cout, 9 * cin) or weight.dtype != torch.bfloat16 or (bias is not None and bias.numel() < cout):
        raise HgError(f"conv3x3_halo: weight {tuple(weight.shape)} != [{cout},{9 * cin}]")
    if out is None:
        out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=x_halo.device)
    lib.check(lib.hg_conv3x3_halo_bf16(_ptr(x_halo), _ptr(weight), _ptr(bias), _ptr(out), _ptr(err_word(x_halo.device)),
                                       _ptr(stats), n, h, w, cin, cout, int(relu), _stream()), "hg_conv3x3_halo_bf16")
    return out


def conv3x3_k3_fusable(n: int, h: int, w: int) -> bool:
    """True when hg_conv3x3_k3_fused_bf16 suits this size (enough 256-position tiles for the CTA pairs, smem budget)."""
    return bool(lib.hg_conv3x3_k3_fusable(n, h, w))


def conv3x3_k3_fused(x_halo: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, w3: torch.Tensor, b3: torch.Tensor, *,
                     n: int, h: int, w: int, residual: Optional[torch.Tensor] = None,
                     up_low: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The tail of a bottleneck in one launch: relu(conv3x3(x_halo) + b2) -> conv1x1 (128 -> 256) + b3 + residual
    (+ nearest-upsampled up_low) -> dense bf16 NHWC [n,h,w,256]   (hg_conv3x3_k3_fused_bf16)."""
    _require_cuda(x_halo, w2, b2, w3, b3, residual, up_low, out)
    if x_halo.numel() != halo_padded_elems(n, h, w, 128) or x_halo.dtype != torch.bfloat16:
        raise HgError("conv3x3_k3_fused: input is not a halo-padded bf16 buffer of [n,h,w,128]")
    if tuple(w2.shape) != (128, 9 * 128) or tuple(w3.shape) != (256, 128) or w2.dtype != torch.bfloat16 or w3.dtype != torch.bfloat16:
        raise HgError(f"conv3x3_k3_fused: weights {tuple(w2.shape)}, {tuple(w3.shape)} != [128,1152], [256,128]")
    if b2.numel() < 128 or b3.numel() < 256 or b2.dtype != torch.float32 or b3.dtype != torch.float32:
        raise HgError("conv3x3_k3_fused: biases must be fp32 [128] and [256]")
    for t, shape, name in ((residual, (n, h, w, 256), "residual"), (up_low, (n, h // 2, w // 2, 256), "up_low"),
                           (out, (n, h, w, 256), "out")):
        if t is not None and (tuple(t.shape) != shape or t.dtype != torch.bfloat16 or not t.is_contiguous()):
            raise HgError(f"conv3x3_k3_fused: {name} must be a contiguous bf16 tensor of shape {shape}")
    if out is None:
        out = torch.empty((n, h, w, 256), dtype=torch.bfloat16, device=x_halo.device)
    lib.check(lib.hg_conv3x3_k3_fused_bf16(_ptr(x_halo), _ptr(w2), _ptr(b2), _ptr(w3), _ptr(b3), _ptr(residual), _ptr(up_low),
                                           _ptr(out), _ptr(err_word(x_halo.device)), n, h, w, _stream()),
              "hg_conv3x3_k3_fused_bf16")
    return out


def stem_packed_buffer(n: int, h: int, w: int, device) -> torch.Tensor:
    """Zero-initialised NHWC4 bf16 staging image with 4 px of horizontal padding on both sides."""
    return torch.zeros((n, h, w + 8, 4), dtype=torch.bfloat16, device=device)


def stem_pack(x_nchw: torch.Tensor, packed: torch.Tensor, flip_w=False) -> torch.Tensor:
    """flip_w: False / True (mirrored), or "both": `packed` is [2n,h,w+8,4], images first, mirrors second, one read."""
    _require_cuda(x_nchw, packed)
    n, c, h, w = x_nchw.shape
    both = flip_w == "both"
    if (c != 3 or x_nchw.dtype != torch.float32 or tuple(packed.shape) != ((2 * n if both else n), h, w + 8, 4)
            or packed.dtype != torch.bfloat16 or not packed.is_contiguous()):
        raise HgError("stem_pack: expects fp32 [n,3,h,w] and a contiguous bf16 [n (2n for both orientations),h,w+8,4] buffer")
    lib.check(lib.hg_stem_pack(_ptr(x_nchw), _ptr(packed), n, h, w, 2 if both else int(bool(flip_w)), _stream()), "hg_stem_pack")
    return packed


def stem_conv(packed: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None):
    """packed [n,h,w+8,4] bf16 -> relu(conv7x7/s2 + bias): bf16 NHWC [n,h/2,w/2,64]."""
    _require_cuda(packed, weight, bias, out)
    n, h, wp, _ = packed.shape
    w = wp - 8
    if tuple(weight.shape) != (64, 224) or weight.dtype != torch.bfloat16 or bias.numel() < 64:
        raise HgError("stem_conv: weight must be bf16 [64,224], bias fp32 [64]")
    if out is None:
        out = torch.empty((n, h // 2, w // 2, 64), dtype=torch.bfloat16, device=packed.device)
    lib.check(lib.hg_stem_conv(_ptr(packed), _ptr(weight), _ptr(bias), _ptr(out), _ptr(err_word(packed.device)), n, h, w,
                               _stream()), "hg_stem_conv")
    return out


def normalize_u8(images_u8: torch.Tensor, mean, std, *, out: Optional[torch.Tensor] = None,
                 packed: Optional[torch.Tensor] = None, flip_w: bool = False) -> torch.Tensor:
    """ToTensor + Normalize on the device (hg_normalize_u8_nhwc): uint8 [n,h,w,3] -> fp32 NCHW [n,3,h,w], bit-identical
    to torch's float32 arithmetic; `packed` additionally receives the stem's NHWC4 bf16 staging image."""
    _require_cuda(images_u8, out, packed)
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3 or not images_u8.is_contiguous():
        raise HgError("normalize_u8: expects a contiguous uint8 [n,h,w,3] tensor")
    n, h, w, _ = images_u8.shape
    if out is None and packed is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=images_u8.device)
    if out is not None and (tuple(out.shape) != (n, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous()):
        raise HgError("normalize_u8: out must be a contiguous fp32 [n,3,h,w] tensor")
    if packed is not None and (tuple(packed.shape) != (n, h, w + 8, 4) or packed.dtype != torch.bfloat16):
        raise HgError("normalize_u8: packed must be a bf16 [n,h,w+8,4] buffer (stem_packed_buffer)")
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    lib.check(lib.hg_normalize_u8_nhwc(_ptr(images_u8), m, sd, _ptr(out), _ptr(packed), n, h, w, int(flip_w), _stream()),
              "hg_normalize_u8_nhwc")
    return out if out is not None else packed


def preprocess_frames_u8(frames_u8: torch.Tensor, mean, std, size, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched Estimator.preprocess_bbox (hg_preprocess_frames_u8): uint8 [n,fh,fw,3] frames -> /255, float64 mean/std
    (None: no normalisation), cv2-style bilinear resize to size=(width, height), fp32 NCHW [n,3,height,width]."""
    _require_cuda(frames_u8, out)
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3 or not frames_u8.is_contiguous():
        raise HgError("preprocess_frames_u8: expects a contiguous uint8 [n,fh,fw,3] tensor")
    n, fh, fw, _ = frames_u8.shape
    w, h = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=frames_u8.device)
    elif tuple(out.shape) != (n, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
        raise HgError("preprocess_frames_u8: out must be a contiguous fp32 [n,3,h,w] tensor")
    if (mean is None) != (std is None):
        raise HgError("preprocess_frames_u8: mean and std go together")
    m = (C.c_double * 3)(*[float(v) for v in mean]) if mean is not None else None
    sd = (C.c_double * 3)(*[float(v) for v in std]) if std is not None else None
    lib.check(lib.hg_preprocess_frames_u8(_ptr(frames_u8), m, sd, _ptr(out), n, fh, fw, h, w, _stream()),
              "hg_preprocess_frames_u8")
    return out


def maxpool2x2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
    lib.check(lib.hg_maxpool2x2_nhwc(_ptr(x), _ptr(out), n, h, w, c, _stream()), "hg_maxpool2x2_nhwc")
    return out


def upsample2x_add(a: torch.Tensor, low: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(a, low, out)
    n, h, w, c = a.shape
    if out is None:
        out = torch.empty_like(a)
    lib.check(lib.hg_upsample2x_add_nhwc(_ptr(a), _ptr(low), _ptr(out), n, h, w, c, _stream()), "hg_upsample2x_add_nhwc")
    return out


def bn_relu(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, scale, shift, out)
    c = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    lib.check(lib.hg_bn_relu_nhwc(_ptr(x), _ptr(scale), _ptr(shift), _ptr(out), x.numel() // c, c, _stream()),
              "hg_bn_relu_nhwc")
    return out


def nchw_to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    lib.check(lib.hg_nchw_f32_to_nhwc_bf16(_ptr(x.float()), _ptr(out), n, c, h, w, _stream()), "hg_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    lib.check(lib.hg_nhwc_bf16_to_nchw_f32(_ptr(x), _ptr(out), n, c, h, w, _stream()), "hg_nhwc_bf16_to_nchw_f32")
    return out


# ------------------------------------------------------------------------------------------------ decode
def _hm(hm: torch.Tensor) -> torch.Tensor:
    if hm.dim() != 4:
        raise AssertionError("Score maps should be 4-dim")       # same message as evaluation.py:13
    if not hm.is_cuda:
        if not torch.cuda.is_available():
            raise HgError("hgb200 decode needs a CUDA device (no CPU fallback)")
        hm = hm.cuda()
    return hm.detach().float().contiguous()


def decode_argmax(hm: torch.Tensor):
    """-> (preds fp32 [b,j,2], maxval fp32 [b,j], idx int32 [b,j]) on the heat map's device."""
    hm = _hm(hm)
    b, j, h, w = hm.shape
    preds = torch.empty((b, j, 2), dtype=torch.float32, device=hm.device)
    maxval = torch.empty((b, j), dtype=torch.float32, device=hm.device)
    idx = torch.empty((b, j), dtype=torch.int32, device=hm.device)
    lib.check(lib.hg_decode_argmax(_ptr(hm), _ptr(preds), _ptr(maxval), _ptr(idx), b, j, h, w, _stream()),
              "hg_decode_argmax")
    return preds, maxval, idx


def decode_final_preds(hm: torch.Tensor, center, scale, output_size) -> torch.Tensor:
    """Batched get_final_preds_v1: center/scale [b,2] (array-like), -> fp64 [b,j,2] on device."""
    hm = _hm(hm)
    b, j, h, w = hm.shape
    c = torch.as_tensor(np.asarray(center, dtype=np.float64).reshape(b, 2)).to(hm.device)
    s = torch.as_tensor(np.asarray(scale, dtype=np.float64).reshape(b, 2)).to(hm.device)
    out = torch.empty((b, j, 2), dtype=torch.float64, device=hm.device)
    lib.check(lib.hg_decode_final_preds(_ptr(hm), _ptr(c), _ptr(s), _ptr(out), b, j, h, w, int(output_size[0]),
                                        int(output_size[1]), _stream()), "hg_decode_final_preds")
    return out


def decode_final_preds_v2(hm: torch.Tensor, center, scale, output_size, refine_joints: int = 2) -> torch.Tensor:
    """Batched get_final_preds_v2 (DARK-style; hg_decode_final_preds_v2): center/scale [b,2] -> fp64 [b,j,2] on device.
    refine_joints=2 is the reference's behaviour (its loop refines joints 0 and 1 only); pass j to refine all."""
    hm = _hm(hm)
    b, j, h, w = hm.shape
    c = torch.as_tensor(np.asarray(center, dtype=np.float64).reshape(b, 2)).to(hm.device)
    s = torch.as_tensor(np.asarray(scale, dtype=np.float64).reshape(b, 2)).to(hm.device)
    out = torch.empty((b, j, 2), dtype=torch.float64, device=hm.device)
    lib.check(lib.hg_decode_final_preds_v2(_ptr(hm), _ptr(c), _ptr(s), _ptr(out), b, j, h, w, int(output_size[0]),
                                           int(output_size[1]), int(refine_joints), _stream()), "hg_decode_final_preds_v2")
    return out


def decode_final_preds_into(hm, center, scale, out, output_size):
    """Pointer-stable form for static plans: every argument is a preallocated device tensor."""
    _require_cuda(hm, center, scale, out)
    b, j, h, w = hm.shape
    lib.check(lib.hg_decode_final_preds(_ptr(hm), _ptr(center), _ptr(scale), _ptr(out), b, j, h, w, int(output_size[0]),
                                        int(output_size[1]), _stream()), "hg_decode_final_preds")
    return out


def flip_average(hm: torch.Tensor, hm_flip: torch.Tensor, perm: torch.Tensor, out: Optional[torch.Tensor] = None):
    if out is None:
        hm, hm_flip = _hm(hm), _hm(hm_flip)
    _require_cuda(hm, hm_flip, perm, out)
    b, j, h, w = hm.shape
    if out is None:
        out = torch.empty_like(hm)
    lib.check(lib.hg_flip_average(_ptr(hm), _ptr(hm_flip), _ptr(perm), _ptr(out), b, j, h, w, _stream()), "hg_flip_average")
    return out


def pck_dists(out_hm: torch.Tensor, tgt_hm: torch.Tensor) -> torch.Tensor:
    out_hm, tgt_hm = _hm(out_hm), _hm(tgt_hm)
    b, j, h, w = out_hm.shape
    d = torch.empty((j, b), dtype=torch.float32, device=out_hm.device)
    lib.check(lib.hg_pck_dists(_ptr(out_hm), _ptr(tgt_hm), _ptr(d), b, j, h, w, _stream()), "hg_pck_dists")
    return d


# ------------------------------------------------------------------------------------------------ loss / targets
_patch_cache = {}


def gaussian_patch(sigma, device) -> torch.Tensor:
    """The reference's un-normalised Gaussian patch (common.py:229-236), computed with numpy in float32
    exactly as the reference does, so on-device targets are bit-identical to generate_target's."""
    key = (float(sigma), str(device))
    p = _patch_cache.get(key)
    if p is None:
        tmp = sigma * 3
        size = 2 * tmp + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, np.newaxis]
        x0 = y0 = size // 2
        g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
        p = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float32)).to(device)
        _patch_cache[key] = p
    return p


def joint_centers(joints: torch.Tensor, vis: torch.Tensor, heatmap_size, image_size, sigma=1):
    """joints/vis fp64 [b,j,3] (device) -> (mu int32 [b,j,2], weight fp32 [b,j])."""
    _require_cuda(joints, vis)
    b, j = joints.shape[:2]
    mu = torch.empty((b, j, 2), dtype=torch.int32, device=joints.device)
    wt = torch.empty((b, j), dtype=torch.float32, device=joints.device)
    lib.check(lib.hg_joint_centers(_ptr(joints), _ptr(vis), _ptr(mu), _ptr(wt), b, j, int(heatmap_size[1]),
                                   int(heatmap_size[0]), int(image_size[0]), int(image_size[1]), int(sigma * 3), _stream()),
              "hg_joint_centers")
    return mu, wt


def gaussian_target(mu: torch.Tensor, weight: torch.Tensor, heatmap_size, sigma=1) -> torch.Tensor:
    _require_cuda(mu, weight)
    b, j = weight.shape
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    patch = gaussian_patch(sigma, mu.device)
    tgt = torch.empty((b, j, h, w), dtype=torch.float32, device=mu.device)
    lib.check(lib.hg_gaussian_target(_ptr(mu), _ptr(weight), _ptr(patch), _ptr(tgt), b, j, h, w, int(sigma * 3), _stream()),
              "hg_gaussian_target")
    return tgt


def jmse_loss(preds: Sequence[torch.Tensor], target: Optional[torch.Tensor], target_weight: Optional[torch.Tensor],
              *, want_grad: bool, grad_scale: float = 1.0, mu: Optional[torch.Tensor] = None, sigma=1):
    """Fused JointsMSE over all stacks.  Returns (loss fp32 [1] device tensor, [grad per stack] or None)."""
    preds = [p if (p.dtype == torch.float32 and p.is_contiguous()) else p.float().contiguous() for p in preds]
    _require_cuda(*preds, target, target_weight, mu)
    b, j, h, w = preds[0].shape
    S = len(preds)
    grads = [torch.empty_like(p) for p in preds] if want_grad else None
    loss = torch.zeros(1, dtype=torch.float32, device=preds[0].device)
    parr = (C.c_void_p * S)(*[p.data_ptr() for p in preds])
    garr = (C.c_void_p * S)(*[g.data_ptr() for g in grads]) if want_grad else None
    tw = target_weight.reshape(b, j).float().contiguous() if target_weight is not None else None
    patch = gaussian_patch(sigma, preds[0].device) if target is None else None
    lib.check(lib.hg_jmse_loss(parr, garr, _ptr(target), _ptr(tw), _ptr(mu), _ptr(patch), int(sigma * 3), _ptr(loss), S,
                               b, j, h, w, C.c_float(grad_scale), _stream()), "hg_jmse_loss")
    return loss, grads


# ------------------------------------------------------------------------------------------------ training
def wgrad(dout: torch.Tensor, z: torch.Tensor, dw: torch.Tensor, *, co_valid: Optional[int] = None, co_first: int = 0,
          ci_valid: Optional[int] = None, taps: int = 1, halo_pitch: int = 0, ld: Optional[int] = None,
          tap_stride: Optional[int] = None, max_ctas: int = 0) -> torch.Tensor:
    """dw (fp32, pre-zeroed or partial) += dout^T . z over rows (hg_wgrad_bf16).

    dout: bf16 [..., co], z: bf16 [..., ci] with the same number of rows (taps=9: both halo-padded flat
    buffers incl. the leading zero row).  dw rows are `ld` floats apart, taps `tap_stride` apart; only dout's
    channels [co_first, co_valid) are written (to dw rows 0..co_valid-co_first-1)."""
    _require_cuda(dout, z, dw)
    if dout.dtype != torch.bfloat16 or z.dtype != torch.bfloat16 or dw.dtype != torch.float32:
        raise HgError("wgrad: dout/z must be bf16, dw fp32")
    co, ci = dout.shape[-1], z.shape[-1]
    rows = dout.numel() // co
    if z.numel() // ci != rows:
        raise HgError(f"wgrad: row counts differ ({rows} vs {z.numel() // ci})")
    ci_valid = ci if ci_valid is None else ci_valid
    co_valid = co if co_valid is None else co_valid
    tap_stride = ci_valid if tap_stride is None else tap_stride
    ld = taps * tap_stride if ld is None else ld
    if dw.numel() < (co_valid - co_first - 1) * ld + (taps - 1) * tap_stride + ci_valid:
        raise HgError("wgrad: dw too small")
    lib.check(lib.hg_wgrad_bf16(_ptr(dout), _ptr(z), _ptr(dw), _ptr(err_word(dout.device)), rows, co, co_first, co_valid, ci,
                                ci_valid, taps, halo_pitch, ld, tap_stride, int(max_ctas), _stream()), "hg_wgrad_bf16")
    return dw


def colreduce_scratch(pixels: int, c: int, device) -> torch.Tensor:
    """Zeroed scratch buffer for the deterministic form of colstats / bn_bwd_reduce (one per launch site)."""
    return torch.zeros(int(lib.hg_colreduce_scratch_bytes(int(pixels), int(c))) // 4, dtype=torch.float32, device=device)


def colstats(x: torch.Tensor, sum_out: torch.Tensor, sumsq_out: Optional[torch.Tensor] = None, c_valid: Optional[int] = None,
             *, shift: bool = False, scratch: Optional[torch.Tensor] = None):
    """sum_out[c] += sum over pixels of x[..., c] (and sumsq_out[c] += sum of squares); x bf16 NHWC.
    shift: sums about x[pixel 0] (see hg_colstats_nhwc); scratch: deterministic fixed-order reduction."""
    _require_cuda(x, sum_out, sumsq_out, scratch)
    c = x.shape[-1]
    lib.check(lib.hg_colstats_nhwc(_ptr(x), _ptr(sum_out), _ptr(sumsq_out), x.numel() // c, c, c if c_valid is None else c_valid,
                                   int(shift), _ptr(scratch), _stream()), "hg_colstats_nhwc")


def bn_train_fwd(x: torch.Tensor, sums: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                 running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                 num_batches_tracked: Optional[torch.Tensor], saved: torch.Tensor, out: torch.Tensor, *, halo: bool = False,
                 relu: bool = True, eps: float = 1e-5, momentum: float = 0.1, shifted: bool = False):
    """Train-mode BatchNorm2d(+ReLU) of a dense bf16 NHWC tensor from batch sums (hg_bn_train_fwd)."""
    _require_cuda(x, sums, gamma, beta, running_mean, running_var, num_batches_tracked, saved, out)
    n, h, w, c = x.shape
    lib.check(lib.hg_bn_train_fwd(_ptr(x), _ptr(sums), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                                  _ptr(num_batches_tracked), _ptr(saved), _ptr(out), n, h, w, c, int(halo), int(relu),
                                  int(shifted), C.c_float(eps), C.c_float(momentum), _stream()), "hg_bn_train_fwd")
    return out


def bn_bwd_reduce(dz: torch.Tensor, x: torch.Tensor, saved: torch.Tensor, sums: torch.Tensor, relu: bool = True,
                  scratch: Optional[torch.Tensor] = None):
    _require_cuda(dz, x, saved, sums, scratch)
    c = x.shape[-1]
    lib.check(lib.hg_bn_bwd_reduce(_ptr(dz), _ptr(x), _ptr(saved), _ptr(sums), x.numel() // c, c, int(relu), _ptr(scratch),
                                   _stream()), "hg_bn_bwd_reduce")


def bn_bwd_apply(dz: torch.Tensor, x: torch.Tensor, saved: torch.Tensor, sums: torch.Tensor, out: torch.Tensor, *,
                 add1: Optional[torch.Tensor] = None, add2: Optional[torch.Tensor] = None,
                 dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None, halo: bool = False,
                 relu: bool = True):
    _require_cuda(dz, x, saved, sums, out, add1, add2, dgamma, dbeta)
    n, h, w, c = x.shape
    lib.check(lib.hg_bn_bwd_apply(_ptr(dz), _ptr(x), _ptr(saved), _ptr(sums), _ptr(add1), _ptr(add2), _ptr(out), _ptr(dgamma),
                                  _ptr(dbeta), n, h, w, c, int(halo), int(relu), _stream()), "hg_bn_bwd_apply")
    return out


def dwconv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], *, relu: bool = False, flip: bool = False,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depthwise 3x3 / pad 1 on NHWC bf16 (hg_dwconv3x3_nhwc); weight fp32 [c*9] in torch's [c,1,3,3] order.
    flip=True applies the reversed taps (the input gradient)."""
    _require_cuda(x, weight, bias, out)
    n, h, w, c = x.shape
    if x.dtype != torch.bfloat16 or weight.dtype != torch.float32 or weight.numel() < c * 9:
        raise HgError("dwconv3x3: x must be bf16 NHWC and weight fp32 with c*9 elements")
    if out is None:
        out = torch.empty_like(x)
    lib.check(lib.hg_dwconv3x3_nhwc(_ptr(x), _ptr(weight), _ptr(bias), _ptr(out), n, h, w, c, int(relu), int(flip), _stream()),
              "hg_dwconv3x3_nhwc")
    return out


def dwconv3x3_wgrad(dout: torch.Tensor, z: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
    """dw (fp32 [c*9], accumulated) += per-channel correlation of dout with the 3x3 neighbourhood of z."""
    _require_cuda(dout, z, dw)
    n, h, w, c = z.shape
    if dout.shape != z.shape or dw.dtype != torch.float32 or dw.numel() < c * 9:
        raise HgError("dwconv3x3_wgrad: shapes differ or dw too small")
    lib.check(lib.hg_dwconv3x3_wgrad(_ptr(dout), _ptr(z), _ptr(dw), n, h, w, c, _stream()), "hg_dwconv3x3_wgrad")
    return dw


def maxpool2x2_bwd(x: torch.Tensor, dpool: torch.Tensor, dx: torch.Tensor, accumulate: bool):
    _require_cuda(x, dpool, dx)
    n, h, w, c = x.shape
    lib.check(lib.hg_maxpool2x2_bwd_nhwc(_ptr(x), _ptr(dpool), _ptr(dx), n, h, w, c, int(accumulate), _stream()),
              "hg_maxpool2x2_bwd_nhwc")
    return dx


def sumpool2x2(dy: torch.Tensor, dlow: torch.Tensor, accumulate: bool = False):
    _require_cuda(dy, dlow)
    n, h, w, c = dy.shape
    lib.check(lib.hg_sumpool2x2_nhwc(_ptr(dy), _ptr(dlow), n, h, w, c, int(accumulate), _stream()), "hg_sumpool2x2_nhwc")
    return dlow


def zero_(t: torch.Tensor):
    """Memset on the current stream (a named call so that launch recorders / the launch DAG see it)."""
    return t.zero_()


def add_inplace(dst: torch.Tensor, src: torch.Tensor):
    _require_cuda(dst, src)
    lib.check(lib.hg_add_inplace_bf16(_ptr(dst), _ptr(src), dst.numel(), _stream()), "hg_add_inplace_bf16")
    return dst


def nchw_to_nhwc_bf16_pad(x: torch.Tensor, out: torch.Tensor):
    """fp32 NCHW [n,c,h,w] -> bf16 NHWC [n,h,w,c_pad] (zero-padded channels)."""
    _require_cuda(x, out)
    n, c, h, w = x.shape
    lib.check(lib.hg_nchw_f32_to_nhwc_bf16_pad(_ptr(x), _ptr(out), n, c, out.shape[-1], h, w, _stream()),
              "hg_nchw_f32_to_nhwc_bf16_pad")
    return out


def make_pack_table(entries, device) -> torch.Tensor:
    """entries: list of dicts(src, src2, dst_f32, dst_fwd, dst_dgrad: tensors or None; co, taps, ci, fwd_ld, fwd_col0,
    dgrad_ld) -> device uint8 tensor holding the hg_pack_entry array."""
    arr = (PackEntry * len(entries))()
    for a, e in zip(arr, entries):
        for k in ("src", "src2", "dst_f32", "dst_fwd", "dst_dgrad"):
            t = e.get(k)
            setattr(a, k, t.data_ptr() if t is not None else None)
        for k in ("co", "taps", "ci", "fwd_ld", "fwd_col0", "dgrad_ld"):
            setattr(a, k, int(e.get(k, 0)))
    raw = bytes(arr)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def pack_weights(table: torch.Tensor, n_entries: int, reads=None, writes=None, which: int = 3):
    """One launch that fills every bf16 GEMM weight layout from the fp32 masters (hg_pack_weights).  `reads` / `writes`
    (lists of tensors) only tell launch recorders which buffers the table points at; the kernel takes the table.
    which: 1 = forward layouts only, 2 = transposed dgrad layouts only, 3 = both."""
    _require_cuda(table)
    lib.check(lib.hg_pack_weights(_ptr(table), n_entries, int(which), _stream()), "hg_pack_weights")


def rmsprop_step(params: torch.Tensor, grads: torch.Tensor, square_avg: torch.Tensor, lr: float, alpha: float = 0.99,
                 eps: float = 1e-8, grad_scale: float = 1.0):
    _require_cuda(params, grads, square_avg)
    lib.check(lib.hg_rmsprop_step(_ptr(params), _ptr(grads), _ptr(square_avg), params.numel(), C.c_float(lr),
                                  C.c_float(alpha), C.c_float(eps), C.c_float(grad_scale), _stream()), "hg_rmsprop_step")


def small_gemm(c: torch.Tensor, a: torch.Tensor, b: torch.Tensor, d: Optional[torch.Tensor], m: int, n: int, k: int,
               sai: int, sak: int, sbk: int, sbj: int, sci: int, scj: int, beta: float = 0.0):
    _require_cuda(c, a, b, d)
    lib.check(lib.hg_small_gemm_f32(_ptr(c), _ptr(a), _ptr(b), _ptr(d), m, n, k, sai, sak, sbk, sbj, sci, scj,
                                    C.c_float(beta), _stream()), "hg_small_gemm_f32")


def jmse_loss_into(preds: Sequence[torch.Tensor], grads: Optional[Sequence[torch.Tensor]], target: Optional[torch.Tensor],
                   target_weight: Optional[torch.Tensor], loss_out: torch.Tensor, *, grad_scale: float = 1.0,
                   mu: Optional[torch.Tensor] = None, sigma=1):
    """Pointer-stable form of jmse_loss for static plans: loss_out (fp32 [1]) must be zeroed by the caller."""
    _require_cuda(*preds, target, target_weight, mu, loss_out)
    b, j, h, w = preds[0].shape
    S = len(preds)
    parr = (C.c_void_p * S)(*[p.data_ptr() for p in preds])
    garr = (C.c_void_p * S)(*[g.data_ptr() for g in grads]) if grads is not None else None
    patch = gaussian_patch(sigma, preds[0].device) if target is None else None
    lib.check(lib.hg_jmse_loss(parr, garr, _ptr(target), _ptr(target_weight), _ptr(mu), _ptr(patch), int(sigma * 3),
                               _ptr(loss_out), S, b, j, h, w, C.c_float(grad_scale), _stream()), "hg_jmse_loss")
