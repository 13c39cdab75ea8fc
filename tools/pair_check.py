"""Paired-CTA 3x3 kernel (HG_CONV3X3_PAIR=1) against the single-CTA kernel and the fp32 reference; timing of both."""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hourglass-pose-estimation_b200"))
from hgb200 import ops
mode = sys.argv[1]
dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
res = {}
for (n, h, w) in [(40, 64, 64), (256, 64, 64), (64, 64, 48), (300, 32, 32), (37, 17, 23)]:
    g = torch.Generator().manual_seed(n + h)
    x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
    wt = torch.randn(128, 128, 3, 3, generator=g) / (3.0 * 128 ** 0.5)
    bias = (torch.randn(128, generator=g) * 0.5).to(dev)
    buf = ops.halo_padded_buffer(n, h, w, 128, dev)
    ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
    wmat = wt.permute(0, 2, 3, 1).reshape(128, 9 * 128).to(torch.bfloat16).to(dev)
    out = ops.conv3x3_halo(buf, wmat, bias, n=n, h=h, w=w, cin=128, cout=128, relu=True)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    if n <= 64:
        ref = F.relu(F.conv2d(x.float().to(dev).permute(0, 3, 1, 2), wt.to(torch.bfloat16).float().to(dev), bias, padding=1))
        err = float((out.float().permute(0, 3, 1, 2) - ref).abs().max())
        print(f"{mode} {(n, h, w)}: max err {err:.4e} (tolerance {float(ref.abs().max()) * 2 ** -7:.4e})")
        assert err <= float(ref.abs().max()) * 2 ** -7
    for _ in range(3):
        ops.conv3x3_halo(buf, wmat, bias, n=n, h=h, w=w, cin=128, cout=128, relu=True, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv3x3_halo(buf, wmat, bias, n=n, h=h, w=w, cin=128, cout=128, relu=True, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{mode} {(n, h, w)}: {us:.1f} us, {2.0 * n * h * w * 1152 * 128 / us / 1e6:.0f} TF/s")
    res[(n, h, w)] = out.cpu()
path = "gpurun_out/pair_check_single.pt"
if mode == "single":
    torch.save(res, path)
else:
    base = torch.load(path)
    for k, v in res.items():
        print(f"pair vs single {k}: bit-identical {torch.equal(v, base[k])}, max diff {float((v.float() - base[k].float()).abs().max()):.3e}")
