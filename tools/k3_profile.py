"""One shape of the fused K2+K3 paired-CTA kernel, a few launches: the target of `ncu --set full --import-source on`.
usage: python tools/k3_profile.py [n h w up]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
n, h, w, use_up = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 64, 64, 0)))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
w2 = (torch.randn(128, 9 * 128, generator=g) / (3.0 * 128 ** 0.5)).to(torch.bfloat16).to(dev)
b2 = (torch.randn(128, generator=g) * 0.5).to(dev)
w3 = (torch.randn(256, 128, generator=g) / 128 ** 0.5).to(torch.bfloat16).to(dev)
b3 = (torch.randn(256, generator=g) * 0.5).to(dev)
res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
up = torch.randn(n, h // 2, w // 2, 256, generator=g).to(torch.bfloat16).to(dev) if use_up else None
buf = ops.halo_padded_buffer(n, h, w, 128, dev)
ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
out = None
for _ in range(4):
    out = ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up, out=out)
torch.cuda.synchronize()
ops.check_err_word(dev)
print("ok", float(out.float().abs().max()))
