"""The bottleneck's opening 1x1 (256 -> 128, bn1+ReLU prologue, halo-padded output) and the other 64x64 1x1 shapes of the
step, a few launches each: timing, and the target of `ncu --set full --import-source on`.
usage: python tools/k1_profile.py [n h w]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
n, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 64, 64)))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
w1 = (torch.randn(128, 256, generator=g) / 16).to(torch.bfloat16).to(dev)
b1 = (torch.randn(128, generator=g) * 0.5).to(dev)
sc, sh = (torch.rand(256, generator=g) + 0.5).to(dev), (torch.randn(256, generator=g) * 0.2).to(dev)
wf = (torch.randn(256, 256, generator=g) / 16).to(torch.bfloat16).to(dev)
bf = (torch.randn(256, generator=g) * 0.5).to(dev)
res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
halo = ops.halo_padded_buffer(n, h, w, 128, dev)
out = torch.empty(n, h, w, 256, dtype=torch.bfloat16, device=dev)
cases = {
    "k1 256->128 pro halo": lambda: ops.conv_nhwc(x, w1, b1, ksize=1, cout=128, relu=True, in_scale=sc, in_shift=sh, out_halo=halo),
    "fc 256->256 relu": lambda: ops.conv_nhwc(x, wf, bf, ksize=1, cout=256, relu=True, out=out),
    "remap 256->256 + res": lambda: ops.conv_nhwc(x, wf, bf, ksize=1, cout=256, residual=res, out=out),
}
px = n * h * w
nbytes = {"k1 256->128 pro halo": px * (512 + 256), "fc 256->256 relu": px * 1024, "remap 256->256 + res": px * 1536}
for name, fn in cases.items():
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name:24s} {us:7.1f} us  {nbytes[name] / us / 1e3:6.0f} GB/s")
ops.check_err_word(dev)
