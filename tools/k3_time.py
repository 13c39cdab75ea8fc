import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
dev = torch.device("cuda")
n, h, w = 256, 64, 64
g = torch.Generator().manual_seed(1)
buf = ops.halo_padded_buffer(n, h, w, 128, dev)
ops.halo_interior(buf, n, h, w, 128).copy_(torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16).to(dev))
w2 = (torch.randn(128, 9 * 128, generator=g) / 34).to(torch.bfloat16).to(dev)
b2 = torch.zeros(128, device=dev)
w3 = (torch.randn(256, 128, generator=g) / 11).to(torch.bfloat16).to(dev)
b3 = torch.zeros(256, device=dev)
res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
out = torch.empty_like(res)
for use_res in (True, False):
    fn = lambda: ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res if use_res else None, out=out)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"HG_K3_DEBUG={os.environ.get('HG_K3_DEBUG', '0')} res={use_res}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us", flush=True)
ops.check_err_word(dev)
