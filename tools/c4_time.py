"""A/B timing of C4-shaped inference (8-stack J=17, 256x192, batch 64 + mirrors)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from src.models import hg
torch.manual_seed(0)
m = hg(num_stacks=8, num_blocks=1, num_classes=17, mobile=False, skip_mode="sum").cuda().eval()
x = torch.randn(128, 3, 256, 192, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = m(x)[-1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = m(x)[-1]
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"C4 inference 128 rows 256x192: {ms:.3f} ms/forward, {128 / ms * 1e3:.0f} rows/s, ragged_halo={'HG_NO_RAGGED_HALO' not in os.environ}, checksum {float(y.float().abs().mean()):.6f}")
if "--breakdown" in sys.argv:
    plan = m.engine(torch.device("cuda")).plan_for(128, 256, 192)
    per = plan.profile(iters=2)
    classes = {}
    for ms_, meta in zip(per, plan.meta):
        c = classes.setdefault(meta["op"], dict(ms=0.0, n=0, flops=meta["flops"], bytes=meta["bytes"]))
        c["ms"] += ms_
        c["n"] += 1
    tot = sum(per)
    print(f"total {tot:.3f} ms over {len(per)} launches")
    for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"])[:28]:
        a = c["ms"] / c["n"]
        print(f"{name:44s} {c['n']:3d} {c['ms']:7.3f} ms  avg {a*1e3:7.1f} us  {c['flops']/(a*1e-3)/1e12:7.1f} TF/s {c['bytes']/(a*1e-3)/1e9:6.0f} GB/s")
