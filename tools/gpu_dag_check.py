"""Race check of the multi-stream launch DAG on the GPU: the same step through a one-stream graph (twice: the
run-to-run noise of the fp32 atomics) and through the K-stream graph; prints relative differences."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hourglass-pose-estimation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import hgb200.train as tr
from src.models import hg
from oracle.hourglass_oracle import make_state_dict
from oracle import train_oracle as T
from oracle.make_golden_inputs import train_inputs

S, J, B, H, W = 2, 16, 4, 128, 128
x, tg, tw = train_inputs(1, B, J, H, W, 1)[0]


def run(k, steps=1, lr=2.5e-4):
    tr.STREAMS = k
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    m = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    m.load_state_dict(sd)
    m = m.cuda().train()
    eng = tr.TrainEngine(m)
    losses = []
    for _ in range(steps):
        losses.append(float(eng.train_step(x.cuda(), tg.cuda(), tw.cuda(), lr)))
    torch.cuda.synchronize()
    tr.ops.check_err_word()
    plan = eng.plans[(B, H, W)]
    return losses, eng.store.G.clone(), [o.clone() for o in plan.outputs], eng.store.P.clone()


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


a = run(1); b = run(1); c = run(6); d = run(8)
for name, r in (("k1 vs k1", b), ("k6 vs k1", c), ("k8 vs k1", d)):
    print(name, "loss", abs(r[0][0] - a[0][0]) / a[0][0], "G", rel(r[1], a[1]), "hm", rel(r[2][-1], a[2][-1]), "P", rel(r[3], a[3]))
sd_ref = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
ref, _, _ = T.train_steps(sd_ref, [(x, tg, tw)] * 6, 2.5e-4)
print("oracle", np.array(ref))
for k in (1, 1, 6, 6):
    print("k", k, np.array(run(k, 6)[0]))
