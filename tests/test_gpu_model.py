"""GPU parity of the whole stacked-hourglass forward (drop-in src.models API) against the live
reference's outputs (golden fixtures), the fp32 oracle and the bf16 numeric-path emulation."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.hourglass_oracle import make_state_dict, hg_forward
from oracle.bf16_emulation import emulate_forward
from oracle import decode_oracle as D

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# north_star tolerance: heat maps within 2e-2 of the peak (bf16 storage, fp32 accumulate)
HEATMAP_TOL = 2e-2


def _build(S, J, nb=1, seed=0, mobile=False, concat=False):
    from src.models import hg
    skip = 'concat' if concat else 'sum'
    sd = make_state_dict(num_stacks=S, num_blocks=nb, num_classes=J, mobile=mobile, skip_mode=skip, seed=seed)
    model = hg(num_stacks=S, num_blocks=nb, num_classes=J, mobile=mobile, skip_mode=skip, out_res=64)
    model.load_state_dict(sd, strict=True)
    return sd, model.to("cuda:0").eval()


def _golden_cases():
    # six configurations run through the live reference (oracle/make_golden.py), incl. mobile=True (depthwise conv2)
    # and skip_mode='concat' 
    return sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz")))


@pytest.mark.parametrize("path", _golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_forward_matches_reference_golden(path):
    z = np.load(path)
    S, J, B, H, W, seed, mobile, concat, nb = [int(v) for v in z["cfg"]]
    sd, model = _build(S, J, nb, seed, bool(mobile), bool(concat))
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed + 1000))
    with torch.no_grad():
        outs = model(x.cuda())
    assert isinstance(outs, list) and len(outs) == S
    emu = emulate_forward(sd, x)
    for i, o in enumerate(outs):
        ref = z[f"out{i}"]
        assert tuple(o.shape) == ref.shape and o.dtype == torch.float32
        o = o.cpu().numpy()
        peak = np.abs(ref).max()
        assert np.abs(o - ref).max() <= HEATMAP_TOL * peak, f"stack {i}: {np.abs(o - ref).max() / peak:.4f} of peak"
        # the CPU emulation rounds at the same points: only accumulation order differs
        e = emu[i].numpy()
        assert np.abs(o - e).max() <= 0.6 * HEATMAP_TOL * peak, f"stack {i} vs emulation: {np.abs(o - e).max() / peak:.4f}"


def test_graph_and_eager_paths_agree_bitwise():
    sd, model = _build(2, 16)
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        model.use_cuda_graph = True
        a = [o.clone() for o in model(x)]
        a2 = [o.clone() for o in model(x)]       # graph replay
        model.use_cuda_graph = False
        b = model(x)
    for u, v, w in zip(a, a2, b):
        assert torch.equal(u, v) and torch.equal(u, w)


def test_flip_forward_equals_forward_of_flipped_input():
    sd, model = _build(2, 16)
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(6)).cuda()
    eng = model.engine()
    a = eng.forward(x, flip=True)
    b = eng.forward(x.flip(-1).contiguous(), flip=False)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_engine_rebuilds_after_weight_update():
    sd, model = _build(1, 16)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(7)).cuda()
    with torch.no_grad():
        a = model(x)[0].clone()
        model.score[0].bias.add_(1.0)
        b = model(x)[0]
    assert torch.allclose(b, a + 1.0, atol=1e-5)


def test_cpu_model_fails_loudly():
    from src.models import hg
    model = hg(num_stacks=1, num_blocks=1, num_classes=16, mobile=False, skip_mode='sum').eval()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))


@pytest.mark.parametrize("J", [21, 14])
def test_c5_four_stack_variants_and_batch_sweep(J):
    """BASELINE config C5: 4-stack hands (21 keypoints) / CrowdPose (14 joints) variants at 256x256, inference at
    several batch sizes incl. 1 and an odd one.  Against the fp32 oracle at the north-star tolerance, and every batch
    size must give the same heat maps for the same image (batch-sharded serving has no cross-image coupling)."""
    sd, model = _build(4, J, seed=11)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(5, 3, 256, 256, generator=g)
    with torch.no_grad():
        ref = hg_forward(sd, x[:2])
        full = [o.cpu() for o in model(x.cuda())]
    assert len(full) == 4 and full[-1].shape == (5, J, 64, 64)
    for o, r in zip(full, ref):
        peak = float(r.abs().max())
        assert float((o[:2] - r).abs().max()) <= HEATMAP_TOL * peak
        same = (o[:2].reshape(2 * J, -1).argmax(1) == r.reshape(2 * J, -1).argmax(1)).float().mean()
        assert float(same) >= 0.7      # random-init maps are flat noise (arg-max ill-conditioned): see test_gpu_trained_accuracy.py
    for b in (1, 2, 3):
        with torch.no_grad():
            part = model(x[:b].cuda())[-1].cpu()
        assert torch.equal(part, full[-1][:b])


def test_empty_batch_gives_empty_heatmaps():
    """Empty input (reference: every torch module passes a 0-sized batch through): a list of S empty heat maps, empty coords."""
    from src.utils.evaluation import get_preds
    _, model = _build(2, 16, seed=1)
    with torch.no_grad():
        out = model(torch.zeros(0, 3, 256, 192, device="cuda"))
    assert len(out) == 2 and all(o.shape == (0, 16, 64, 48) and o.dtype == torch.float32 and o.is_cuda for o in out)
    assert get_preds(out[-1]).shape == (0, 16, 2)


def test_maximum_batch_crosses_the_32_bit_element_limit():
    """Maximum sizes: 2112 rows at 64x64x256 channels are 2.2 G elements / 4.4 GB per activation tensor -- past both 2^31
    elements and 2^32 bytes, so any 32-bit index arithmetic in a kernel shows up.  No oracle can run this batch; the
    check is the size-independent property of the path: an image's heat maps do not depend on its batch (bit-identical
    to the same image in a batch of three), checked at the start, the middle and the far end of the large batch."""
    sd, model = _build(1, 16, seed=5)
    g = torch.Generator().manual_seed(6)
    x3 = torch.randn(3, 3, 256, 256, generator=g)
    n = 2112
    with torch.no_grad():
        small = model(x3.cuda())[-1]
        big_in = x3.repeat(n // 3, 1, 1, 1).cuda()
        big = model(big_in)[-1]
    torch.cuda.synchronize()
    from hgb200 import ops
    ops.check_err_word(torch.device("cuda:0"))
    assert big.shape == (n, 16, 64, 64)
    for start in (0, 1056, n - 3):
        assert torch.equal(big[start:start + 3], small), start
    assert torch.equal(big.view(n // 3, 3, 16, 64, 64)[::37], small.expand(len(range(0, n // 3, 37)), -1, -1, -1, -1))
    del big, big_in
    model._engine = None
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n,h,w,cin,res,up,x2", [(2, 64, 64, 128, True, False, 0), (3, 32, 32, 128, True, False, 0),
                                                (5, 16, 16, 128, True, True, 0), (9, 8, 8, 64, False, False, 64),
                                                (4, 4, 4, 128, True, False, 0), (6, 2, 2, 64, False, False, 0),
                                                (1, 64, 64, 128, True, True, 0)])
def test_conv1x1_fused_maxpool_output(n, h, w, cin, res, up, x2):
    """hg_conv_desc.pool_out: the 2x2 max-pool written by the 1x1 GEMM's epilogue equals the pool kernel applied to the
    GEMM's own (bf16) result, bit for bit, and the main result is untouched."""
    from hgb200 import ops
    g = torch.Generator().manual_seed(h * 7 + cin)
    cout = 256
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).cuda()
    k = cin + x2
    wt = (torch.randn(cout, k, generator=g) / k ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    r = torch.randn(n, h, w, cout, generator=g).to(torch.bfloat16).cuda() if res else None
    lo = torch.randn(n, h // 2, w // 2, cout, generator=g).to(torch.bfloat16).cuda() if up else None
    xx = torch.randn(n, h, w, x2, generator=g).to(torch.bfloat16).cuda() if x2 else None
    assert ops.conv_pool_fusable(h, w, cout, k) and not ops.conv_pool_fusable(h, w, cout, 256)
    plain = ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, residual=r, up_low=lo, x2=xx)
    pooled = torch.full((n, h // 2, w // 2, cout), 7.0, dtype=torch.bfloat16, device="cuda")
    fused = ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, residual=r, up_low=lo, x2=xx, pool_out=pooled)
    torch.cuda.synchronize()
    ops.check_err_word()
    assert torch.equal(fused, plain)
    assert torch.equal(pooled, ops.maxpool2x2(plain))


def test_pool_fusion_does_not_change_the_network(monkeypatch):
    import hgb200.engine as E
    sd, model = _build(2, 16)
    x = torch.randn(3, 3, 128, 128, generator=torch.Generator().manual_seed(8)).cuda()
    outs = {}
    monkeypatch.setattr(E, "POOL_IN", False)          # the producer-side fusion alone (the consumer-side one: next test)
    for fuse in (True, False):
        monkeypatch.setattr(E, "FUSE_POOL", fuse)
        eng = E.HourglassEngine(sd, "cuda:0")
        plan = eng.build_plan(3, 128, 128, use_graph=True)
        plan.input.copy_(x)
        plan.run()
        torch.cuda.synchronize()
        outs[fuse] = ([o.clone() for o in plan.outputs], plan.num_launches)
    # three pools per stack (32^2, 16^2, 8^2 inputs: K3 of a bottleneck, K = 128) ride in their producers' epilogues, plus
    # the first stack's 64^2 input (layer3's K3); the later 64^2 inputs come from the K = 256 remap GEMM and keep the kernel
    assert outs[True][1] == outs[False][1] - (2 * 3 + 1)
    for a, b in zip(outs[True][0], outs[False][0]):
        assert torch.equal(a, b)



def test_pool_from_the_up1_prologue_does_not_change_the_network(monkeypatch):
    """hg_conv_desc.pool_in (engine.POOL_IN): a level's pooled input comes out of the prologue of the 1x1 conv that opens the
    level's up1 bottleneck, which is emitted BEFORE the lower pyramid instead of after it.  Same heat maps bit for bit, one
    launch fewer per level that takes the halo path (here 32^2 and 16^2: 4 rows of 128^2 input, two stacks)."""
    import hgb200.engine as E
    sd, model = _build(2, 16)
    x = torch.randn(4, 3, 128, 128, generator=torch.Generator().manual_seed(9)).cuda()
    outs = {}
    monkeypatch.setattr(E, "FUSE_POOL", False)
    for pool_in in (True, False):
        monkeypatch.setattr(E, "POOL_IN", pool_in)
        eng = E.HourglassEngine(sd, "cuda:0")
        plan = eng.build_plan(4, 128, 128, use_graph=True)
        plan.input.copy_(x)
        plan.run()
        torch.cuda.synchronize()
        ops_ = [m["op"] for m in plan.meta]
        outs[pool_in] = ([o.clone() for o in plan.outputs], plan.num_launches, sum(o.endswith("_poolin") for o in ops_),
                         sum(o.startswith("maxpool") for o in ops_))
    assert outs[True][2] == 4 and outs[False][2] == 0
    assert outs[True][1] == outs[False][1] - 4 and outs[True][3] == outs[False][3] - 4
    for a, b in zip(outs[True][0], outs[False][0]):
        assert torch.equal(a, b)
