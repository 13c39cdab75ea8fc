"""In-graph kernel timeline of the C2 flip-test step (CUPTI through torch.profiler: kernel start / end stamps of graph
replays, which neither eager event timing nor ncu -- it serialises -- can give): per-kernel-name time inside the graph, the
idle gaps between consecutive kernels, and the step's busy fraction.  usage: python tools/graph_trace.py [infer|train]"""
import os, sys, json, collections
import numpy as np
import torch
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "hourglass-pose-estimation_b200"))
from torch.profiler import profile, ProfilerActivity
what = sys.argv[1] if len(sys.argv) > 1 else "infer"
dev = torch.device("cuda")
if what == "infer":
    from bench import build_model
    from hgb200.infer import FlipTestPipeline
    B = 128
    pipe = FlipTestPipeline(build_model(dev).engine(dev), B, 256, 256)
    pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
    x = torch.randn(B, 3, 256, 256, device=dev)
    run = lambda: pipe.infer_device(x)
else:
    from hgb200 import ops
    from hgb200.train import train_engine
    from src.models import hg
    B, J = 32, 16
    torch.manual_seed(0)
    eng = train_engine(hg(num_stacks=8, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum", out_res=64).to(dev).train())
    x = torch.randn(B, 3, 256, 256, device=dev)
    tgt = torch.rand(B, J, 64, 64, device=dev)
    wt = torch.ones(B, J, device=dev)
    run = lambda: eng.train_step(x, tgt, wt, 2.5e-4)
for _ in range(5):
    run()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        run()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
per = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][-60:]
    per[k][0] += 1
    per[k][1] += e.time_range.end - e.time_range.start
# union of busy intervals (kernels overlap under PDL / several streams)
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, t
    else:
        cur_e = max(cur_e, t)
busy += cur_e - cur_s
span = t1 - t0
print(f"{what}: {len(ev)} kernel records over {N} steps; span {span / N / 1e3:.3f} ms/step, some kernel running {busy / N / 1e3:.3f} ms/step "
      f"({busy / span:.3f}), sum of kernel durations {sum(v[1] for v in per.values()) / N / 1e3:.3f} ms/step")
for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us / N / 1e3:8.3f} ms/step  {n // N:5d} launches/step  avg {us / n:7.1f} us  {k}")
