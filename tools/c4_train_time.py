"""C4-shaped training step (8-stack J=17, 256x192, batch 64): step time and per-class breakdown."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
from hgb200.train import train_engine
from src.models import hg
B, J, H, W, lr = int(os.environ.get("C4_BATCH", "64")), 17, 256, 192, 2.5e-4
torch.manual_seed(0)
model = hg(num_stacks=8, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum").cuda().train()
eng = train_engine(model)
rng = np.random.RandomState(100)
x = torch.randn(B, 3, H, W, device="cuda")
joints = np.zeros((B, J, 3)); joints[..., 0], joints[..., 1] = rng.uniform(0, W, (B, J)), rng.uniform(0, H, (B, J))
vis = (rng.rand(B, J, 1) < 0.8).astype(np.float64).repeat(3, 2)
jt, vs = torch.from_numpy(joints).cuda(), torch.from_numpy(vis).cuda()
def step():
    mu, wt = ops.joint_centers(jt, vs, (W // 4, H // 4), (W, H), 1)
    tgt = ops.gaussian_target(mu, wt, (W // 4, H // 4), 1)
    return eng.train_step(x, tgt, wt, lr)
for _ in range(4):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = step()
e1.record(); torch.cuda.synchronize()
ops.check_err_word(torch.device("cuda"))
ms = e0.elapsed_time(e1) / 10
print(f"C4 train batch {B} 256x192 J=17: {ms:.3f} ms/step, {B / ms * 1e3:.0f} img/s, loss {float(loss):.5f}")
plan = eng.plans[(B, H, W)]
per = plan.profile(iters=2)
classes = {}
for ms_, meta in zip(per, plan.meta):
    c = classes.setdefault(meta["op"], dict(ms=0.0, n=0, flops=meta["flops"], bytes=meta["bytes"]))
    c["ms"] += ms_; c["n"] += 1
print(f"sum of launches {sum(per):.3f} ms over {len(per)}")
for name, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"])[:32]:
    a = c["ms"] / c["n"]
    print(f"{name:44s} {c['n']:3d} {c['ms']:7.3f} ms  avg {a*1e3:7.1f} us  {c['flops']/(a*1e-3)/1e12:7.1f} TF/s {c['bytes']/(a*1e-3)/1e9:6.0f} GB/s")
