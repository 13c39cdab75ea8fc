"""CPU emulation (plain torch) of the libhgb200 entry points hgb200.train drives, with the same argument
conventions and the same bf16 rounding points.  TEST INFRASTRUCTURE ONLY: it lets the host-side logic of
the training plan (gradient routing, buffer aliasing, weight packing, the remap chain rule) be checked
against the oracle without a GPU.  The product never imports this file."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from hgb200.ops import halo_interior, halo_padded_elems, gaussian_patch  # noqa: F401  (pure torch)

BF = torch.bfloat16          # storage type of activations; tests may switch it to float32 for exact comparisons


def require_device(device):
    """The emulation runs anywhere (the real hgb200.ops.require_device raises off-CUDA)."""


def halo_padded_buffer(n, h, w, c, device):
    return torch.zeros(halo_padded_elems(n, h, w, c), dtype=BF, device=device)


def check_err_word(device=None):
    pass


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _store(out, y_nchw):
    out.copy_(y_nchw.permute(0, 2, 3, 1).to(out.dtype))
    return out


def conv_nhwc(x, weight, bias, *, ksize, cout, relu=False, in_scale=None, in_shift=None, residual=None, up_low=None,
              x2=None, out=None, out_nchw_f32=None, heads=False, out_halo=None, stats=None, pool_out=None, pool_in=None):
    n, h, w, cin = x.shape
    if pool_in is not None:               # the 2x2 max-pool of the RAW input (before the prologue)
        _store(pool_in, F.max_pool2d(_nchw(x), 2, 2))
    taps = ksize * ksize
    wm = weight.float()[:cout]
    w4 = wm[:, :taps * cin].reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    b = bias[:cout].float() if bias is not None else None
    y = F.conv2d(_nchw(x), w4, b, padding=ksize // 2)
    if x2 is not None:
        y = y + F.conv2d(_nchw(x2), wm[:, taps * cin:, None, None])
    if residual is not None:
        y = y + _nchw(residual)
    if up_low is not None:
        y = y + F.interpolate(_nchw(up_low), scale_factor=2, mode="nearest")
    if relu:
        y = F.relu(y)
    if stats is not None:                 # the epilogue's fused per-channel sums of the fp32 result (before rounding)
        stats[:cout] += y.sum((0, 2, 3))
        stats[cout:2 * cout] += (y * y).sum((0, 2, 3))
    if heads:
        out_nchw_f32.copy_(y)
        return out_nchw_f32
    _store(out, y)
    if pool_out is not None:              # the 2x2 max-pool of the STORED (rounded) result
        _store(pool_out, F.max_pool2d(_nchw(out), 2, 2))
    return out


def conv_pool_fusable(h, w, cout, k):
    return cout == 256 and k <= 128 and 2 <= w <= 64 and (w & (w - 1)) == 0 and h % 2 == 0 and 128 % (2 * w) == 0


def conv3x3_halo(x_halo, weight, bias, *, n, h, w, cin, cout, relu=False, out=None, stats=None):
    x = halo_interior(x_halo, n, h, w, cin)
    return conv_nhwc(x, weight, bias, ksize=3, cout=cout, relu=relu, out=out, stats=stats)


def dwconv3x3(x, weight, bias, *, relu=False, flip=False, out=None):
    c = x.shape[-1]
    w4 = weight[:c * 9].float().reshape(c, 1, 3, 3)
    if flip:
        w4 = w4.flip(2, 3)
    y = F.conv2d(_nchw(x), w4, bias[:c].float() if bias is not None else None, padding=1, groups=c)
    if relu:
        y = F.relu(y)
    if out is None:
        out = torch.empty_like(x)
    return _store(out, y)


def dwconv3x3_wgrad(dout, z, dw):
    c = z.shape[-1]
    g, zz = _nchw(dout), F.pad(_nchw(z), (1, 1, 1, 1))
    h, w = g.shape[2], g.shape[3]
    acc = torch.stack([(g * zz[:, :, ky:ky + h, kx:kx + w]).sum((0, 2, 3)) for ky in range(3) for kx in range(3)], 1)
    dw[:c * 9] += acc.reshape(-1)
    return dw


def stem_im2col(x_nchw, flip_w=False, out=None):
    n, c, h, w = x_nchw.shape
    cols = F.unfold(x_nchw, 7, padding=3, stride=2)              # [n, c*49, L], row index = c*49 + ky*7 + kx
    cols = cols.view(n, 3, 49, -1).permute(0, 3, 2, 1).reshape(n, h // 2, w // 2, 147)    # k = tap*3 + c
    out.zero_()
    out[..., :147] = cols.to(out.dtype)
    return out


def maxpool2x2(x, out=None):
    return _store(out, F.max_pool2d(_nchw(x), 2, 2))


def maxpool2x2_bwd(x, dpool, dx, accumulate):
    with torch.enable_grad():          # the plan may be running inside an autograd.Function (grad mode off there)
        xf = _nchw(x).detach().requires_grad_(True)
        F.max_pool2d(xf, 2, 2).backward(_nchw(dpool))
    g = xf.grad
    if accumulate:
        g = g + _nchw(dx)
    return _store(dx, g)


def sumpool2x2(dy, dlow, accumulate=False):
    g = F.avg_pool2d(_nchw(dy), 2) * 4
    if accumulate:
        g = g + _nchw(dlow)
    return _store(dlow, g)


def zero_(t):
    return t.zero_()


def add_inplace(dst, src):
    dst.copy_((dst.float() + src.float()).to(dst.dtype))
    return dst


def nchw_to_nhwc_bf16_pad(x, out):
    out.zero_()
    out[..., :x.shape[1]] = x.permute(0, 2, 3, 1).to(out.dtype)
    return out


def colreduce_scratch(pixels, c, device):
    return torch.zeros(16, dtype=torch.float32, device=device)


def colstats(x, sum_out, sumsq_out=None, c_valid=None, *, shift=False, scratch=None):
    c = x.shape[-1]
    cv = c if c_valid is None else c_valid
    xf = x.float().reshape(-1, c)
    if shift:
        xf = xf - xf[0]            # sums about x[pixel 0] (hg_colstats_nhwc)
    sum_out[:cv] += xf.sum(0)[:cv]
    if sumsq_out is not None:
        sumsq_out[:cv] += (xf * xf).sum(0)[:cv]


def bn_train_fwd(x, sums, gamma, beta, running_mean, running_var, num_batches_tracked, saved, out, *, halo=False,
                 relu=True, eps=1e-5, momentum=0.1, shifted=False):
    n, h, w, c = x.shape
    N = n * h * w
    mean = sums[:c] / N
    var = (sums[c:2 * c] / N - mean * mean).clamp_min(0)
    if shifted:
        mean = mean + x.float().reshape(-1, c)[0]
    invstd = torch.rsqrt(var + eps)
    sc = gamma * invstd
    sh = beta - mean * sc
    saved[:c], saved[c:2 * c], saved[2 * c:3 * c], saved[3 * c:4 * c] = mean, invstd, sc, sh
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * var * N / max(N - 1, 1))
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    y = x.float() * sc + sh
    if relu:
        y = F.relu(y)
    dst = halo_interior(out, n, h, w, c) if halo else out
    dst.copy_(y.to(BF))
    return out


def _masked(dz, x, saved, relu):
    c = x.shape[-1]
    sc, sh = saved[2 * c:3 * c], saved[3 * c:4 * c]
    dy = dz.float()
    if relu:
        dy = dy * ((x.float() * sc + sh) > 0)
    xhat = (x.float() - saved[:c]) * saved[c:2 * c]
    return dy, xhat


def bn_bwd_reduce(dz, x, saved, sums, relu=True, scratch=None):
    c = x.shape[-1]
    dy, xhat = _masked(dz, x, saved, relu)
    sums[:c] += dy.reshape(-1, c).sum(0)
    sums[c:2 * c] += (dy * xhat).reshape(-1, c).sum(0)


def bn_bwd_apply(dz, x, saved, sums, out, *, add1=None, add2=None, dgamma=None, dbeta=None, halo=False, relu=True):
    n, h, w, c = x.shape
    N = n * h * w
    dy, xhat = _masked(dz, x, saved, relu)
    r = saved[2 * c:3 * c] * (dy - sums[:c] / N - xhat * sums[c:2 * c] / N)
    if add1 is not None:
        r = r + add1.float()
    if add2 is not None:
        r = r + add2.float()
    if dgamma is not None:
        dgamma.copy_(sums[c:2 * c])
        dbeta.copy_(sums[:c])
    dst = halo_interior(out, n, h, w, c) if halo else out
    dst.copy_(r.to(BF))
    return out


def wgrad(dout, z, dw, *, co_valid=None, co_first=0, ci_valid=None, taps=1, halo_pitch=0, ld=None, tap_stride=None,
          max_ctas=0):
    co, ci = dout.shape[-1], z.shape[-1]
    cov = co if co_valid is None else co_valid
    civ = ci if ci_valid is None else ci_valid
    tap_stride = civ if tap_stride is None else tap_stride
    ld = taps * tap_stride if ld is None else ld
    a = dout.float().reshape(-1, co)
    b = z.float().reshape(-1, ci)
    rows = a.shape[0]
    for tap in range(taps):
        off = 0 if taps == 1 else (tap // 3 - 1) * halo_pitch + tap % 3 - 1
        bs = torch.zeros_like(b)
        lo, hi = max(0, -off), min(rows, rows - off)
        bs[lo:hi] = b[lo + off:hi + off]
        g = a.t() @ bs                                          # [co, ci]
        view = torch.as_strided(dw, (cov - co_first, civ), (ld, 1), dw.storage_offset() + tap * tap_stride)
        view += g[co_first:cov, :civ]
    return dw


def make_pack_table(entries, device):
    return entries


def pack_weights(table, n_entries, reads=None, writes=None, which=3):
    for e in table:
        co, taps, ci = e["co"], e["taps"], e["ci"]
        v = e["src"][:co * taps * ci].float().clone()
        if e.get("src2") is not None:
            v = v + e["src2"][:co * taps * ci]
        if e.get("dst_f32") is not None and which & 1:
            e["dst_f32"][:co * taps * ci] = v
        w = v.view(co, taps, ci)
        if e.get("dst_fwd") is not None and which & 1:
            e["dst_fwd"][:co, e["fwd_col0"]:e["fwd_col0"] + taps * ci] = w.reshape(co, -1).to(BF)
        if e.get("dst_dgrad") is not None and which & 2:
            e["dst_dgrad"][:, :taps * co] = w.flip(1).permute(2, 1, 0).reshape(ci, taps * co).to(BF)


def rmsprop_step(params, grads, square_avg, lr, alpha=0.99, eps=1e-8, grad_scale=1.0):
    g = grads * grad_scale
    square_avg.mul_(alpha).addcmul_(g, g, value=1 - alpha)
    params.sub_(lr * g / (square_avg.sqrt() + eps))


def small_gemm(c, a, b, d, m, n, k, sai, sak, sbk, sbj, sci, scj, beta=0.0):
    A = torch.as_strided(a, (m, k), (sai, sak), a.storage_offset())
    Bm = torch.as_strided(b, (k, n), (sbk, sbj), b.storage_offset())
    Cv = torch.as_strided(c, (m, n), (sci, scj), c.storage_offset())
    r = A @ Bm
    if beta != 0.0:
        r = r + beta * Cv
    if d is not None:
        r = r + torch.as_strided(d, (m, n), (sci, scj), d.storage_offset())
    Cv.copy_(r)


def jmse_loss_into(preds, grads, target, target_weight, loss_out, *, grad_scale=1.0, mu=None, sigma=1):
    b, j, h, w = preds[0].shape
    wt = target_weight.reshape(b, j, 1, 1) if target_weight is not None else torch.ones(b, j, 1, 1)
    for i, p in enumerate(preds):
        d = (p - target) * wt
        loss_out += 0.5 * (d * d).sum() / (b * j * h * w)
        if grads is not None:
            grads[i].copy_(wt * wt * (p - target) / (b * j * h * w) * grad_scale)
