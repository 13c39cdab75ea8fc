"""Verbose GPU sweep (not a pytest file): runs every conv case and the small kernels, keeps going after
failures and prints where the errors are.  `python tools/gpu_debug.py [filter]` under gpurun."""
import os
import sys
import time
import traceback

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200"))

from tests.gpu_cases import CONV_CASES, make_conv_case, conv_reference, run_conv_case, no_tf32  # noqa: E402


def describe_mismatch(out, ref, tol):
    bad = (out - ref).abs() > tol
    n, c, h, w = ref.shape
    print(f"    bad elements: {int(bad.sum())} / {bad.numel()}")
    if bad.any():
        idx = bad.nonzero()
        print("    first bad (n,c,y,x):", idx[:6].tolist())
        print("    bad per channel-block of 8:", bad.sum(dim=(0, 2, 3)).reshape(-1, 8).sum(1).tolist()[:32])
        print("    bad per row y:", bad.sum(dim=(0, 1, 3)).tolist()[:64])
        print("    bad per col x:", bad.sum(dim=(0, 1, 2)).tolist()[:64])
        print("    bad per image:", bad.sum(dim=(1, 2, 3)).tolist()[:32])
        i = tuple(idx[0].tolist())
        print("    sample out/ref:", float(out[i]), float(ref[i]))


def main():
    no_tf32()
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    fails = 0
    for case in CONV_CASES:
        if filt and filt not in case[0]:
            continue
        try:
            t = make_conv_case(case, dev)
            ref = conv_reference(t)
            t0 = time.time()
            out = run_conv_case(t)
            dt = time.time() - t0
            scale = float(ref.abs().max())
            tol = scale * (2 ** -7) if not t["heads"] else scale * 1e-3
            err = float((out - ref).abs().max())
            ok = err <= tol and bool(torch.isfinite(out).all())
            print(f"[{'ok' if ok else 'FAIL'}] {case[0]:24s} err={err:.4g} tol={tol:.4g} scale={scale:.3g} ({dt*1e3:.1f} ms)")
            if not ok:
                fails += 1
                describe_mismatch(out, ref, tol)
        except Exception:
            fails += 1
            print(f"[EXC] {case[0]}")
            traceback.print_exc()
    print("conv failures:", fails)
    return fails


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
