"""The reference-facing Python API (src.utils / src.loss) on the GPU, against golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from oracle import golden_inputs as GI

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_get_preds_cpu_and_cuda_tensor_inputs():
    from src.utils.evaluation import get_preds
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    hm = GI.heatmap_cases()[6]
    p_cpu = get_preds(torch.from_numpy(hm))                 # CPU tensor in -> CPU tensor out (kernel still runs on the GPU)
    assert not p_cpu.is_cuda and p_cpu.dtype == torch.float32
    np.testing.assert_array_equal(p_cpu.numpy(), z["preds6"])
    p_gpu = get_preds(torch.from_numpy(hm).cuda())
    assert p_gpu.is_cuda
    np.testing.assert_array_equal(p_gpu.cpu().numpy(), z["preds6"])
    with pytest.raises(AssertionError):
        get_preds(torch.zeros(3, 4, 5))


def test_accuracy_matches_reference():
    from src.utils.evaluation import accuracy
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    for i, (pred, tgt) in enumerate(GI.accuracy_cases()):
        p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda()
        np.testing.assert_allclose(accuracy(p, t, None, 0.5), z[f"acc{i}"], atol=1e-12)
        np.testing.assert_allclose(accuracy(p[:, [1, 3, 5]], t[:, [1, 3, 5]], [1, 3, 5], 0.5), z[f"acc_sub{i}"], atol=1e-12)
        np.testing.assert_allclose(accuracy(p, t, None, 0.2), z[f"acc_thr{i}"], atol=1e-12)


def test_get_final_preds_v1_and_transforms():
    from src.utils.inference import get_final_preds_v1
    from src.utils.transforms import get_affine_transform, transform_preds, fliplr_joints
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    cases = GI.heatmap_cases()
    centers, scales = GI.decode_args(cases)
    for i in (0, 2, 6, 7):
        hm = cases[i]
        B, J, H, W = hm.shape
        for b in range(B):
            out = get_final_preds_v1(torch.from_numpy(hm[b:b + 1]), centers[i][b], scales[i][b], (W, H))
            assert out.dtype == np.float64 and out.shape == (J, 2)
            np.testing.assert_allclose(out, z[f"final{i}"][b], rtol=0, atol=1e-9)
    for a, m in zip(GI.affine_args(), z["affine_mats"]):
        M = get_affine_transform(np.array(a[0:2]), np.array(a[2:4]), 0, (int(a[4]), int(a[5])), inv=1)
        np.testing.assert_allclose(M, m, rtol=0, atol=1e-9)
    fj, fv = fliplr_joints(z["flip_j"].copy(), z["flip_v"].copy(), 256, D.MPII_FLIP_PAIRS)
    np.testing.assert_array_equal(fj, z["flip_j_out"])
    np.testing.assert_array_equal(fv, z["flip_v_out"])
    c = np.array([[3.0, 4.0], [10.25, 7.0]])
    np.testing.assert_allclose(transform_preds(c, [100, 50], [1.5, 2.5], (64, 48)),
                               D.transform_preds(c, [100, 50], [1.5, 2.5], (64, 48)), atol=1e-9)


def test_mse_loss_module_forward_backward():
    from src.loss.mse import MSELoss, JointsMSELossOnTheFly
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    stride = int(z["grad_stride"])
    for i, c in enumerate(GI.loss_cases()):
        tg, tw = z[f"target{i}"], z[f"tw{i}"]
        outs = [torch.from_numpy(tg + n).cuda().requires_grad_(True) for n in c["noise"]]
        loss = MSELoss(True)(outs, torch.from_numpy(tg).cuda(), torch.from_numpy(tw).cuda())
        assert loss.dim() == 0
        assert abs(float(loss) - float(z[f"loss{i}"])) <= 2e-5 * float(z[f"loss{i}"])
        (2.0 * loss).backward()
        for s, o in enumerate(outs):
            np.testing.assert_allclose(o.grad.cpu().numpy().reshape(-1)[::stride], 2.0 * z[f"grad{i}_{s}"], rtol=2e-5,
                                       atol=1e-10)
        # on-the-fly targets from joints: same loss without a target tensor
        outs2 = [torch.from_numpy(tg + n).cuda() for n in c["noise"]]
        crit = JointsMSELossOnTheFly(c["isz"], c["hsz"], 1)
        l2 = crit(outs2, torch.from_numpy(c["joints"]), torch.from_numpy(c["vis"]))
        assert abs(float(l2) - float(z[f"loss{i}"])) <= 2e-5 * float(z[f"loss{i}"])
    l_nw = MSELoss(False)([torch.from_numpy(z["target0"] + n).cuda() for n in GI.loss_cases()[0]["noise"]],
                          torch.from_numpy(z["target0"]).cuda(), torch.from_numpy(z["tw0"]).cuda())
    assert abs(float(l_nw) - float(z["loss_nw0"])) <= 2e-5 * float(z["loss_nw0"])


def test_flip_test_pipeline_matches_oracle_on_its_own_heatmaps():
    from src.models import hg
    from hgb200.infer import FlipTestPipeline
    from oracle.hourglass_oracle import make_state_dict
    sd = make_state_dict(num_stacks=2, num_classes=16, seed=0)
    model = hg(num_stacks=2, num_blocks=1, num_classes=16, mobile=False, skip_mode='sum')
    model.load_state_dict(sd)
    model = model.cuda().eval()
    x = torch.randn(3, 3, 128, 128, generator=torch.Generator().manual_seed(9))
    eng = model.engine()
    pipe = FlipTestPipeline(eng, 3, 128, 128)
    centers = np.array([[60.0, 70.0], [64.0, 64.0], [10.0, 100.0]])
    scales = np.array([[0.64, 0.64], [1.0, 1.0], [0.3, 0.5]])
    pipe.set_affine(centers, scales)
    coords = pipe.infer_device(x.cuda()).cpu().numpy()
    # the two halves of the doubled batch equal separate plain / mirrored forwards
    plain = eng.forward(x.cuda(), flip=False)[-1].cpu().numpy()
    mirrored = eng.forward(x.cuda(), flip=True)[-1].cpu().numpy()
    avg = D.flip_average(plain, mirrored, D.MPII_FLIP_PAIRS)
    hm = pipe.plan.heatmap.cpu().numpy()
    np.testing.assert_allclose(hm, avg, rtol=0, atol=1e-6)
    np.testing.assert_allclose(coords, D.get_final_preds_batch(hm, centers, scales, (32, 32)), rtol=0, atol=1e-9)
    # host pipeline returns the same coordinates
    hb = [x.pin_memory(), x.pin_memory(), x.pin_memory()]
    outs = list(pipe.infer_host(hb))
    assert len(outs) == 3
    for o in outs:
        np.testing.assert_array_equal(o, coords)


def test_device_prefetcher_and_lagged_scalar():
    """hgb200/prefetch.py: batches arrive in order and intact although the next copy is already in flight while the
    consumer still works on the current buffers; the lagged scalar reader returns every step's value, one step late."""
    from hgb200.prefetch import DevicePrefetcher, LaggedScalar
    host = [(torch.full((64, 1024), float(i)).pin_memory(), torch.full((7,), i, dtype=torch.int64).pin_memory())
            for i in range(9)]
    reader = LaggedScalar("cuda")
    seen, lagged = [], []
    for i, (a, b) in enumerate(DevicePrefetcher(iter(host), "cuda")):
        assert a.is_cuda and b.is_cuda
        for _ in range(20):                       # keep the consumer busy on this buffer while the next copy runs
            a = a * 1.0
        s = (a.sum() / a.numel()) + b[0].float() * 100
        seen.append(s)
        v = reader.push(s)
        if v is not None:
            lagged.append(v)
    lagged.append(reader.flush())
    torch.cuda.synchronize()
    want = [i + 100.0 * i for i in range(9)]
    assert [float(s) for s in seen] == want
    assert lagged == want
