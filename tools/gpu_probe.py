"""Hardware probe: UMMA smem descriptors with row-shifted (non-1024-aligned) start addresses."""
import ctypes as C
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200"))
from hgb200._lib import lib_path  # noqa: E402


def main():
    dll = C.CDLL(lib_path())
    g = torch.Generator().manual_seed(0)
    a = torch.randn(192, 64, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
    shifts = [0, 1, 2, 3, 5, 7, 8, 9, 16, 33, 64]
    out = torch.zeros(len(shifts), 2, 128, 64, device="cuda")
    sh = (C.c_int32 * len(shifts))(*shifts)
    rc = dll.hg_debug_shifted_desc(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), sh,
                                   len(shifts), None)
    torch.cuda.synchronize()
    print("rc", rc)
    for i, s in enumerate(shifts):
        ref = a[s:s + 128].float() @ b.float().t()
        for v in range(2):
            err = float((out[i, v] - ref).abs().max())
            print(f"shift {s:3d} variant {'base_offset' if v else 'plain      '}: max err {err:.4g}  {'OK' if err < 1e-2 else 'WRONG'}")


if __name__ == "__main__":
    main()
