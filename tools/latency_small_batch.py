"""Small-batch latency of the headline model (8-stack J=16, 256x256): the reference's scripts/estimate.py case (one frame,
no flip test) and the flip-test pipeline at batch 1 / 2 / 8.  CUDA events around graph replays, 50 each."""
import os, sys
import numpy as np, torch
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "hourglass-pose-estimation_b200"))
from bench import build_model
from hgb200 import ops
from hgb200.infer import FlipTestPipeline
dev = torch.device("cuda")
model = build_model(dev)
eng = model.engine(dev)
def timed(fn, n=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x1 = torch.randn(1, 3, 256, 256, device=dev)
with torch.no_grad():
    print(f"model(x)[-1], batch 1, no flip (estimate.py): {timed(lambda: model(x1)[-1]):.3f} ms")
for B in (1, 2, 8):
    pipe = FlipTestPipeline(eng, B, 256, 256)
    pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
    x = torch.randn(B, 3, 256, 256, device=dev)
    ms = timed(lambda: pipe.infer_device(x))
    print(f"flip-test pipeline (2 forwards + flip average + decode), batch {B}: {ms:.3f} ms = {B / ms * 1e3:.0f} img/s, "
          f"{pipe.plan.streams} streams, {pipe.launches_per_batch} launches")
ops.check_err_word(dev)
