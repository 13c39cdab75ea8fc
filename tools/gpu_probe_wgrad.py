"""Verbose probe of hg_wgrad_bf16 (MN-major UMMA descriptors) against a torch fp32 reference.
Keeps going after a failure so one gpurun call yields the full picture.  Not a pytest file."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200")]
from hgb200 import ops  # noqa: E402


def case_1x1(rows, co, ci, seed=0, co_valid=None):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(rows, co, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(rows, ci, generator=g).to(torch.bfloat16).cuda()
    cv = co_valid or co
    dw = torch.zeros(cv, ci, device="cuda")
    ops.wgrad(a, b, dw, co_valid=cv)
    torch.cuda.synchronize()
    ops.check_err_word()
    ref = a.float().t()[:cv] @ b.float()
    err = (dw - ref).abs().max().item() / ref.abs().max().item()
    return err


def case_3x3(n, h, w, co, ci, seed=0):
    g = torch.Generator().manual_seed(seed)
    P = w + 1
    d = torch.randn(n, h, w, co, generator=g).to(torch.bfloat16)
    z = torch.randn(n, h, w, ci, generator=g).to(torch.bfloat16)
    dh = ops.halo_padded_buffer(n, h, w, co, "cuda")
    zh = ops.halo_padded_buffer(n, h, w, ci, "cuda")
    ops.halo_interior(dh, n, h, w, co).copy_(d.cuda())
    ops.halo_interior(zh, n, h, w, ci).copy_(z.cuda())
    dw = torch.zeros(co, 9, ci, device="cuda")
    ops.wgrad(dh.view(-1, co), zh.view(-1, ci), dw, taps=9, halo_pitch=P)
    torch.cuda.synchronize()
    ops.check_err_word()
    # reference: conv weight gradient, [co, ci, 3, 3] -> [co, tap, ci]
    zz = z.float().permute(0, 3, 1, 2).cuda().requires_grad_(False)
    wt = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    torch.backends.cudnn.allow_tf32 = False
    y = torch.nn.functional.conv2d(zz, wt, padding=1)
    y.backward(d.float().permute(0, 3, 1, 2).cuda())
    ref = wt.grad.permute(0, 2, 3, 1).reshape(co, 9, ci)
    err = (dw - ref).abs().max().item() / ref.abs().max().item()
    return err


def main():
    print("swap knob:", os.environ.get("HG_WGRAD_SWAP_LBO_SBO"))
    for (rows, co, ci, cv) in [(64, 128, 64, None), (128, 128, 128, None), (640, 256, 128, None), (4096, 128, 256, None),
                               (5000, 256, 256, None), (70000, 256, 256, None), (333, 64, 64, None), (1000, 64, 256, 16),
                               (8192, 64, 192, None)]:
        try:
            e = case_1x1(rows, co, ci, co_valid=cv)
            print(f"1x1 rows={rows} co={co} ci={ci} cv={cv}: rel err {e:.3e} {'OK' if e < 2e-3 else 'FAIL'}")
        except Exception as ex:  # noqa: BLE001
            print(f"1x1 rows={rows} co={co} ci={ci}: EXC {ex}")
    for (n, h, w, co, ci) in [(1, 8, 8, 128, 128), (2, 16, 16, 128, 128), (2, 64, 64, 128, 128), (3, 4, 4, 128, 128),
                              (1, 64, 48, 128, 128), (1, 128, 128, 64, 64), (5, 7, 3, 64, 64)]:
        try:
            e = case_3x3(n, h, w, co, ci)
            print(f"3x3 n={n} {h}x{w} co={co} ci={ci}: rel err {e:.3e} {'OK' if e < 2e-3 else 'FAIL'}")
        except Exception as ex:  # noqa: BLE001
            print(f"3x3 n={n} {h}x{w}: EXC {ex}")


if __name__ == "__main__":
    main()
