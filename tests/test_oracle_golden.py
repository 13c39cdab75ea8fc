"""Pin the CPU oracle to the live reference's outputs (tests/golden/*.npz).

The fixtures were produced by oracle/make_golden.py importing /root/reference.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from oracle import loss_oracle as L
from oracle import golden_inputs as GI
from oracle.hourglass_oracle import make_state_dict, hg_forward, conv_flops_per_image

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_model_oracle_matches_reference(path):
    z = np.load(path)
    S, J, B, H, W, seed, mobile, concat, nb = [int(v) for v in z["cfg"]]
    sd = make_state_dict(num_stacks=S, num_blocks=nb, num_classes=J, mobile=bool(mobile),
                         skip_mode="concat" if concat else "sum", seed=seed)
    assert len(sd) == int(z["n_keys"])
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed + 1000))
    with torch.no_grad():
        outs = hg_forward(sd, x)
    assert len(outs) == S
    for i, o in enumerate(outs):
        ref = z[f"out{i}"]
        assert o.shape == ref.shape
        # same fp32 ATen kernels in a different association order -> tiny drift only
        np.testing.assert_allclose(o.numpy(), ref, rtol=2e-4, atol=2e-4 * np.abs(ref).max())


def test_flop_count_matches_survey():
    sd = make_state_dict(num_stacks=2, num_classes=16)
    assert abs(conv_flops_per_image(sd) / 1e9 - 17.978) < 0.01


def test_get_preds_and_final_preds_bit_exact():
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    cases = GI.heatmap_cases()
    centers, scales = GI.decode_args(cases)
    assert len(cases) == int(z["n"])
    for i, hm in enumerate(cases):
        np.testing.assert_array_equal(D.get_preds(hm), z[f"preds{i}"])
        B, J, H, W = hm.shape
        fin = D.get_final_preds_batch(hm, centers[i], scales[i], (W, H))
        np.testing.assert_allclose(fin, z[f"final{i}"], rtol=0, atol=1e-9)


def test_quirk_table():
    # SURVEY.md A8 probe: peaks at (0,0) and (3,2) on a 4x5 map -> [[5,0],[3,3]]
    hm = GI.heatmap_cases()[0]
    np.testing.assert_array_equal(D.get_preds(hm)[0], np.array([[5, 0], [3, 3]], np.float32))


def test_affine_matrix():
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    args = GI.affine_args()
    for a, m in zip(args, z["affine_mats"]):
        M = D.affine_inv_matrix(a[0:2], a[2:4], (int(a[4]), int(a[5])))
        np.testing.assert_allclose(M, m, rtol=0, atol=1e-9)
    # closed form of SURVEY.md A10
    np.testing.assert_allclose(D.affine_inv_matrix([100, 50], [1.5, 2.5], (64, 48)),
                               [[4.6875, 0, -50], [0, 4.6875, -62.5]], atol=1e-9)


def test_accuracy_pck():
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    for i, (pred, tgt) in enumerate(GI.accuracy_cases()):
        np.testing.assert_allclose(D.accuracy(pred, tgt, None, 0.5), z[f"acc{i}"], atol=1e-12)
        np.testing.assert_allclose(D.accuracy(pred[:, [1, 3, 5]], tgt[:, [1, 3, 5]], [1, 3, 5], 0.5),
                                   z[f"acc_sub{i}"], atol=1e-12)
        np.testing.assert_allclose(D.accuracy(pred, tgt, None, 0.2), z[f"acc_thr{i}"], atol=1e-12)


def test_fliplr_joints():
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    fj, fv = D.fliplr_joints(z["flip_j"], z["flip_v"], 256, D.MPII_FLIP_PAIRS)
    np.testing.assert_array_equal(fj, z["flip_j_out"])
    np.testing.assert_array_equal(fv, z["flip_v_out"])


def test_generate_target_and_loss():
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    stride = int(z["grad_stride"])
    for i, c in enumerate(GI.loss_cases()):
        tg, tw = L.generate_target_batch(c["joints"], c["vis"], c["isz"], c["hsz"], 1)
        np.testing.assert_array_equal(tg, z[f"target{i}"])
        np.testing.assert_array_equal(tw, z[f"tw{i}"])
        outs = [tg + n for n in c["noise"]]
        loss, grads = L.joints_mse(outs, tg, tw, True)
        assert abs(loss - float(z[f"loss{i}"])) <= 2e-6 * abs(loss)
        loss_nw, _ = L.joints_mse(outs, tg, tw, False)
        assert abs(loss_nw - float(z[f"loss_nw{i}"])) <= 2e-6 * abs(loss_nw)
        for s, g in enumerate(grads):
            ref = z[f"grad{i}_{s}"]
            np.testing.assert_allclose(g.reshape(-1)[::stride], ref, rtol=1e-5, atol=1e-12)
    # edge cases called out in SURVEY.md 8c
    c = GI.loss_cases()[0]
    tg, tw = L.generate_target_batch(c["joints"][:1], c["vis"][:1], c["isz"], c["hsz"], 1)
    assert tw[0, 0, 0] == 0 and tg[0, 0].sum() == 0            # off-map joint
    assert tg[0, 1].max() == 1.0 and tg[0, 1, 0, 0] == 1.0     # corner joint, clipped patch


def test_flip_average_definition():
    rng = np.random.RandomState(0)
    hm = rng.rand(2, 16, 8, 8).astype(np.float32)
    perm = D.flip_perm(16, D.MPII_FLIP_PAIRS)
    # a model that is exactly flip-equivariant returns hm itself after flip-back
    hm_f = hm[:, perm][:, :, :, ::-1]
    np.testing.assert_array_equal(D.flip_average(hm, hm_f, D.MPII_FLIP_PAIRS), hm)


def test_preprocess_oracle_matches_reference():
    """oracle/preprocess_oracle.py against Estimator.preprocess_bbox and ToTensor+Normalize run live
    (oracle/make_golden_preprocess.py).  ToTensor+Normalize is float32 arithmetic: bit-exact.  preprocess_bbox goes
    through cv2.resize in float64 whose summation order OpenCV does not document: the restatement agrees to 1e-15 in
    float64, i.e. after the reference's own cast to float32 it is identical up to a rare last-bit flip (<= 1 ulp)."""
    from oracle import preprocess_oracle as P
    z = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    for i in range(int(z["n_crops"])):
        got = P.normalize_u8(z[f"crop{i}"], z["crop_mean"], z["crop_std"])
        assert got.dtype == np.float32 and np.array_equal(got, z[f"crop_out{i}"])
    for i in range(int(z["n_frames"])):
        dataset, res = str(z[f"frame_cfg{i}"][0]), int(z[f"frame_cfg{i}"][1])
        got = P.preprocess_bbox(z[f"frame{i}"], dataset, (res, res))
        ref = z[f"frame_out{i}"]
        assert got.shape == ref.shape == (1, 3, res, res) and got.dtype == np.float32
        ulp = np.spacing(np.abs(ref).astype(np.float32))
        assert np.all(np.abs(got - ref) <= ulp)
        assert np.mean(got != ref) < 1e-3


def test_dark_decode_oracle_matches_reference():
    """get_final_preds_v2 restatement (oracle/decode_oracle.py) against the live reference (oracle/make_golden_dark.py).
    The blur's float64 summation order inside OpenCV is undocumented but the reference stores the blurred map into a
    float32 array, which erases the difference: identical on every fixture up to the affine solve's 1e-14."""
    from oracle.golden_inputs import dark_cases
    z = np.load(os.path.join(GOLDEN, "dark.npz"))
    cases = dark_cases()
    assert int(z["n"]) == len(cases)
    with np.errstate(invalid="ignore"):
        for i, c in enumerate(cases):
            got = D.get_final_preds_v2(c["hm"], c["center"], c["scale"], c["output_size"])
            np.testing.assert_allclose(got, z[f"pred{i}"], rtol=0, atol=1e-9)
