"""8-stack parity debug: per-stack heat-map error of model.forward, engine.forward(flip), and the flip-test pipeline."""
import os, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200")]
from oracle.hourglass_oracle import make_state_dict, hg_forward, calibrate_bn
from oracle import decode_oracle as D
from hgb200.infer import FlipTestPipeline
from hgb200.flip import MPII_FLIP_PAIRS
from src.models import hg
S, J, H, W = 8, 16, 256, 256
Bs = [int(a) for a in sys.argv[1:]] or [2, 8]
for B in Bs:
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 3, H, W, generator=g)
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum", out_res=64)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    with torch.no_grad():
        ref = [r.numpy() for r in hg_forward(sd, x)]
        reff = [r.numpy() for r in hg_forward(sd, x.flip(-1))]
        outs = [o.cpu().numpy() for o in model(x.cuda())]
        eng = model.engine()
        outf = [o.cpu().numpy() for o in eng.forward(x.cuda(), flip=True)]
    e = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    print(f"B={B} forward      :", " ".join(f"{e(o, r):.4f}" for o, r in zip(outs, ref)), flush=True)
    print(f"B={B} forward flip :", " ".join(f"{e(o, r):.4f}" for o, r in zip(outf, reff)), flush=True)
    pipe = FlipTestPipeline(eng, B, H, W, flip_pairs=MPII_FLIP_PAIRS)
    pipe.set_affine(np.tile([[128.0, 128.0]], (B, 1)), np.tile([[1.28, 1.28]], (B, 1)))
    pipe.infer_device(x.cuda())
    hm = pipe.plan.heatmap.cpu().numpy()
    last = pipe.plan.outputs[-1].cpu().numpy()
    print(f"B={B} pipeline last-stack rows [0,B) vs ref {e(last[:B], ref[-1]):.4f}; rows [B,2B) vs ref(flipped input) {e(last[B:], reff[-1]):.4f}")
    avg = D.flip_average(ref[-1], reff[-1], MPII_FLIP_PAIRS)
    print(f"B={B} pipeline flip-average vs oracle {e(hm, avg):.4f}; flip_average kernel on its own inputs vs numpy {e(hm, D.flip_average(last[:B], last[B:], MPII_FLIP_PAIRS)):.6f}")
