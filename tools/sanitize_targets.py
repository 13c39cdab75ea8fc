"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): the paired-CTA K2+K3 kernel (all four
residual / upsample instantiations) against the two-kernel path, the 1x1 kernel with the bn1 prologue and halo store, the
single-CTA halo 3x3, the deterministic per-channel reductions, and one whole eager training step of a 1-stack network.
    compute-sanitizer --tool memcheck  python tools/sanitize_targets.py [k3|conv|train|all]
    compute-sanitizer --tool racecheck python tools/sanitize_targets.py ..."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "hourglass-pose-estimation_b200")]
from hgb200 import ops  # noqa: E402

dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def k3():
    for (n, h, w, use_res, use_up) in [(5, 64, 64, True, False), (5, 64, 64, True, True), (5, 64, 64, False, False),
                                       (5, 64, 64, False, True), (20, 32, 32, True, True)]:
        g = torch.Generator().manual_seed(n + h)
        x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
        w2 = (torch.randn(128, 9 * 128, generator=g) / (3.0 * 128 ** 0.5)).to(torch.bfloat16).to(dev)
        b2 = (torch.randn(128, generator=g) * 0.5).to(dev)
        w3 = (torch.randn(256, 128, generator=g) / 128 ** 0.5).to(torch.bfloat16).to(dev)
        b3 = (torch.randn(256, generator=g) * 0.5).to(dev)
        res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev) if use_res else None
        up = torch.randn(n, h // 2, w // 2, 256, generator=g).to(torch.bfloat16).to(dev) if use_up else None
        buf = ops.halo_padded_buffer(n, h, w, 128, dev)
        ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
        assert ops.conv3x3_k3_fusable(n, h, w)
        z2 = ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True)
        want = ops.conv_nhwc(z2, w3, b3, ksize=1, cout=256, residual=res, up_low=up)
        got = ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up)
        torch.cuda.synchronize()
        ops.check_err_word(dev)
        print(f"k3 {(n, h, w, use_res, use_up)}: bit-identical {torch.equal(got, want)}", flush=True)
        assert torch.equal(got, want)


def conv():
    g = torch.Generator().manual_seed(3)
    n, h, w = 3, 64, 64
    x = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev)
    w1 = (torch.randn(128, 256, generator=g) / 16).to(torch.bfloat16).to(dev)
    b1 = torch.randn(128, generator=g).to(dev)
    s1, t1 = (0.5 + torch.rand(256, generator=g)).to(dev), torch.randn(256, generator=g).to(dev)
    halo = ops.halo_padded_buffer(n, h, w, 128, dev)
    ops.conv_nhwc(x, w1, b1, ksize=1, cout=128, relu=True, in_scale=s1, in_shift=t1, out_halo=halo)
    dense = ops.conv_nhwc(x, w1, b1, ksize=1, cout=128, relu=True, in_scale=s1, in_shift=t1)
    torch.cuda.synchronize()
    assert torch.equal(ops.halo_interior(halo, n, h, w, 128), dense)
    pooled = torch.empty(n, h // 2, w // 2, 256, dtype=torch.bfloat16, device=dev)
    w3 = (torch.randn(256, 128, generator=g) / 11).to(torch.bfloat16).to(dev)
    b3 = torch.randn(256, generator=g).to(dev)
    out = ops.conv_nhwc(dense, w3, b3, ksize=1, cout=256, residual=x, pool_out=pooled)
    torch.cuda.synchronize()
    assert torch.equal(pooled, ops.maxpool2x2(out))
    sums = torch.zeros(512, device=dev)
    scr = ops.colreduce_scratch(n * h * w, 256, dev)
    ops.colstats(out, sums[:256], sums[256:], shift=True, scratch=scr)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    print("conv1x1 prologue / halo / pool / fixed-order colstats ok", flush=True)


def train():
    import hgb200.train as tr
    from src.models import hg
    tr.STREAMS = 1
    tr.DETERMINISTIC = True
    torch.manual_seed(0)
    model = hg(num_stacks=1, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum").to(dev).train()
    eng = tr.TrainEngine(model)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 64, 64, generator=g).to(dev)
    tgt = torch.rand(2, 16, 16, 16, generator=g).to(dev)
    tw = torch.ones(2, 16, device=dev)
    loss = eng.train_step(x, tgt, tw, 2.5e-4, use_graph=False)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    print(f"train step ok, loss {float(loss):.5f}", flush=True)


if what in ("k3", "all"):
    k3()
if what in ("conv", "all"):
    conv()
if what in ("train", "all"):
    train()
