"""Parity on the HEADLINE configuration itself (BASELINE.json configs[1], what bench.py times): 8-stack hourglass, J = 16,
256x256 inputs, through `FlipTestPipeline` -- the benched path, CUDA graph, paired-CTA K2+K3 kernel, flip average and
decode -- against the fp32 oracle (reference: src/models/hourglass.py:69-90 is where the error compounds over the stacks).

North-star tolerances, written here: heat maps within 2e-2 of the peak (every stack, and the flip-averaged last stack the
decode reads); decode bit-exact / 1e-9 on identical heat maps."""
import numpy as np
import pytest
import torch

from oracle.hourglass_oracle import make_state_dict, hg_forward, calibrate_bn
from oracle import decode_oracle as D

pytestmark = pytest.mark.gpu

HEATMAP_TOL = 2e-2
B, H, W, S, J = 8, 256, 256, 8, 16


def _model(sd):
    from src.models import hg
    model = hg(num_stacks=S, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum", out_res=64)
    model.load_state_dict(sd, strict=True)
    return model.to("cuda:0").eval()


@pytest.mark.parametrize("calibrated", [False, True], ids=["randomised_bn_as_benched", "calibrated_bn"])
def test_c2_eight_stack_flip_pipeline_matches_oracle(calibrated):
    from hgb200 import ops
    from hgb200.infer import FlipTestPipeline
    from hgb200.flip import MPII_FLIP_PAIRS
    sd = make_state_dict(num_stacks=S, num_blocks=1, num_classes=J, seed=0)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 3, H, W, generator=g)
    if calibrated:
        # unit-variance pre-activations, as in a trained network (oracle/hourglass_oracle.py:calibrate_bn)
        calibrate_bn(sd, torch.randn(4, 3, H, W, generator=g))
    model = _model(sd)
    pipe = FlipTestPipeline(model.engine(), B, H, W, flip_pairs=MPII_FLIP_PAIRS)
    fused = [m["op"] for m in pipe.plan.meta if m["op"].startswith("conv3x3h_k3_fused")]
    assert len(fused) >= 8 * 7, "the benched plan runs the bottleneck tails on the paired-CTA K2+K3 kernel"
    centers = np.tile([[128.0, 128.0]], (B, 1))
    scales = np.tile([[1.28, 1.28]], (B, 1))
    pipe.set_affine(centers, scales)
    coords = pipe.infer_device(x.cuda()).cpu().numpy()
    hm = pipe.plan.heatmap.cpu().numpy()
    ops.check_err_word()
    with torch.no_grad():
        ref = hg_forward(sd, x)
        ref_flip = hg_forward(sd, x.flip(-1))
    ref_avg = D.flip_average(ref[-1].numpy(), ref_flip[-1].numpy(), MPII_FLIP_PAIRS)
    peak = np.abs(ref_avg).max()
    err = np.abs(hm - ref_avg).max() / peak
    assert err <= HEATMAP_TOL, f"flip-averaged stack-8 heat map: {err:.4f} of peak"
    # decode of the SAME heat maps: get_final_preds_v1 for every image, bit-exact arg-max, 1e-9 on the fp64 affine
    np.testing.assert_allclose(coords, D.get_final_preds_batch(hm, centers, scales, (W // 4, H // 4)), rtol=0, atol=1e-9)
    preds, _, _ = ops.decode_argmax(pipe.plan.heatmap)
    np.testing.assert_array_equal(preds.cpu().numpy(), D.get_preds(hm))
    # every stack's heat map through the module's own forward (all 8 heads, no flip)
    with torch.no_grad():
        outs = [o.cpu().numpy() for o in model(x.cuda())]
    assert len(outs) == S
    errs = []
    for o, r in zip(outs, ref):
        r = r.numpy()
        errs.append(np.abs(o - r).max() / np.abs(r).max())
    print(f"\n8-stack {'calibrated' if calibrated else 'randomised-BN'}: per-stack max error of peak "
          + " ".join(f"{e:.4f}" for e in errs) + f"; flip-averaged last stack {err:.4f}")
    assert max(errs) <= HEATMAP_TOL, errs
