"""Verbose model-level sweep: error per stack vs fp32 oracle / emulation, and rough timing."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200"))

from oracle.hourglass_oracle import make_state_dict, hg_forward  # noqa: E402
from oracle.bf16_emulation import emulate_forward  # noqa: E402
from oracle.decode_oracle import get_preds  # noqa: E402


def main():
    from src.models import hg
    torch.set_num_threads(os.cpu_count())
    for (S, J, B, H, W) in [(2, 16, 2, 256, 256), (8, 16, 2, 256, 256), (2, 17, 2, 256, 192)]:
        sd = make_state_dict(num_stacks=S, num_classes=J, seed=0)
        model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode='sum')
        model.load_state_dict(sd)
        model = model.cuda().eval()
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(2))
        with torch.no_grad():
            ref = hg_forward(sd, x)
            emu = emulate_forward(sd, x)
            out = [o.cpu() for o in model(x.cuda())]
        for s in sorted(set([0, S // 2, S - 1])):
            r, e, o = ref[s].numpy(), emu[s].numpy(), out[s].numpy()
            peak = np.abs(r).max()
            agree = (get_preds(r) == get_preds(o)).all(axis=2).mean()
            print(f"S={S} J={J} {H}x{W} stack{s}: peak={peak:.3g} gpu-vs-fp32={np.abs(o-r).max()/peak:.4f} "
                  f"gpu-vs-emu={np.abs(o-e).max()/peak:.4f} emu-vs-fp32={np.abs(e-r).max()/peak:.4f} argmax-agree={agree:.3f}")
    # timing, 8-stack, several batch sizes
    sd = make_state_dict(num_stacks=8, num_classes=16, seed=0)
    model = hg(num_stacks=8, num_blocks=1, num_classes=16, mobile=False, skip_mode='sum')
    model.load_state_dict(sd)
    model = model.cuda().eval()
    eng = model.engine()
    for B in (1, 8, 32, 128):
        x = torch.randn(B, 3, 256, 256, device="cuda")
        plan = eng.plan_for(B, 256, 256, False, True)
        plan.input.copy_(x)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 56.189e9 * B / (ms * 1e-3) / 1e12
        print(f"B={B}: {ms:.3f} ms/forward  {B/ms*1e3:.0f} img/s  {tf:.1f} TFLOP/s ({tf/1616.4*100:.1f}% of measured bf16 peak)"
              f"  launches={plan.num_launches} arena={plan.arena_bytes/2**20:.0f} MiB")


if __name__ == "__main__":
    main()
