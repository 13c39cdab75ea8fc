"""Fused K2+K3 paired-CTA kernel against the two-kernel path (bit level) and timing of both."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
dev = torch.device("cuda")
for (n, h, w, use_res, use_up) in [(40, 64, 64, True, False), (40, 64, 64, True, True), (64, 64, 48, True, True), (256, 64, 64, True, False),
                                   (256, 64, 64, True, True), (256, 32, 32, True, True), (300, 16, 16, True, False), (41, 30, 22, False, False)]:
    g = torch.Generator().manual_seed(n + h)
    x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(128, 9 * 128, generator=g) / (3.0 * 128 ** 0.5)).to(torch.bfloat16).to(dev)
    b2 = (torch.randn(128, generator=g) * 0.5).to(dev)
    w3 = (torch.randn(256, 128, generator=g) / 128 ** 0.5).to(torch.bfloat16).to(dev)
    b3 = (torch.randn(256, generator=g) * 0.5).to(dev)
    res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev) if use_res else None
    up = torch.randn(n, h // 2, w // 2, 256, generator=g).to(torch.bfloat16).to(dev) if use_up else None
    buf = ops.halo_padded_buffer(n, h, w, 128, dev)
    ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
    print(f"{(n, h, w, use_res, use_up)} fusable={ops.conv3x3_k3_fusable(n, h, w)}", flush=True)
    z2 = ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True)
    want = ops.conv_nhwc(z2, w3, b3, ksize=1, cout=256, residual=res, up_low=up)
    got = ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    d = (got.float() - want.float()).abs()
    print(f"   bit-identical {torch.equal(got, want)}  max diff {float(d.max()):.3e}  scale {float(want.float().abs().max()):.3f}  mismatching {int((d > 0).sum())}", flush=True)
    def timeit(fn):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    t_two = timeit(lambda: ops.conv_nhwc(ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True, out=z2), w3, b3, ksize=1, cout=256, residual=res, up_low=up, out=want))
    t_one = timeit(lambda: ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up, out=got))
    print(f"   two kernels {t_two:.1f} us, fused {t_one:.1f} us", flush=True)
