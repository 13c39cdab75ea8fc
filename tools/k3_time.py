"""Timing of the fused K2+K3 kernel alone on the bench shapes (A/B runs: tools/k3_ab.sh swaps the library)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
dev = torch.device("cuda")
out = []
for (n, h, w, use_up) in [(256, 64, 64, False), (256, 64, 64, True), (256, 32, 32, False), (256, 32, 32, True), (256, 16, 16, False)]:
    g = torch.Generator().manual_seed(n + h)
    x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(128, 9 * 128, generator=g) / (3.0 * 128 ** 0.5)).to(torch.bfloat16).to(dev)
    b2 = (torch.randn(128, generator=g) * 0.5).to(dev)
    w3 = (torch.randn(256, 128, generator=g) / 128 ** 0.5).to(torch.bfloat16).to(dev)
    b3 = (torch.randn(256, generator=g) * 0.5).to(dev)
    res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
    up = torch.randn(n, h // 2, w // 2, 256, generator=g).to(torch.bfloat16).to(dev) if use_up else None
    buf = ops.halo_padded_buffer(n, h, w, 128, dev)
    ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
    got = ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up)
    for _ in range(5):
        ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up, out=got)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up, out=got)
    e1.record(); torch.cuda.synchronize()
    ops.check_err_word(dev)
    out.append(f"{h}x{w}{'+up' if use_up else ''} {e0.elapsed_time(e1) / 40 * 1e3:.1f}")
print(" | ".join(out), "| checksum", float(got.float().sum()))
