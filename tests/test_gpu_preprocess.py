"""GPU parity of the on-device input preprocessing through the C ABI, against the CPU
oracle (bit-exact: same float32 / float64 arithmetic) and the live reference's outputs (tests/golden/preprocess.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_normalize_u8_matches_totensor_normalize_bit_exactly():
    from hgb200 import ops
    z = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    for i in range(int(z["n_crops"])):
        crop = torch.from_numpy(z[f"crop{i}"]).cuda()
        got = ops.normalize_u8(crop[None], z["crop_mean"], z["crop_std"])
        assert got.shape == (1, 3) + crop.shape[:2]
        assert np.array_equal(got[0].cpu().numpy(), z[f"crop_out{i}"])           # the live reference
    # a batch, every byte value, both outputs at once; the packed image is the stem's own layout
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (3, 64, 128, 3), generator=g, dtype=torch.uint8)
    x[0, 0, :, 0] = torch.arange(128, dtype=torch.uint8) * 2
    mean, std = [0.4003, 0.4314, 0.4534], [0.2466, 0.2467, 0.2562]
    packed = ops.stem_packed_buffer(3, 64, 128, "cuda")
    out = torch.empty(3, 3, 64, 128, device="cuda")
    ops.normalize_u8(x.cuda(), mean, std, out=out, packed=packed)
    ref = np.stack([P.normalize_u8(x[b].numpy(), mean, std) for b in range(3)])
    assert np.array_equal(out.cpu().numpy(), ref)
    want = ops.stem_packed_buffer(3, 64, 128, "cuda")
    ops.stem_pack(out, want)
    assert torch.equal(packed, want)
    packed_f = ops.stem_packed_buffer(3, 64, 128, "cuda")
    ops.normalize_u8(x.cuda(), mean, std, packed=packed_f, flip_w=True)
    want_f = ops.stem_packed_buffer(3, 64, 128, "cuda")
    ops.stem_pack(out, want_f, flip_w=True)
    assert torch.equal(packed_f, want_f)


def test_preprocess_frames_matches_oracle_and_reference():
    from hgb200 import ops
    z = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    for i in range(int(z["n_frames"])):
        dataset, res = str(z[f"frame_cfg{i}"][0]), int(z[f"frame_cfg{i}"][1])
        frame = z[f"frame{i}"]
        ms = P.dataset_mean_std(dataset)
        mean, std = ms if ms is not None else (None, None)
        got = ops.preprocess_frames_u8(torch.from_numpy(frame[None]).cuda(), mean, std, (res, res)).cpu().numpy()
        assert np.array_equal(got, P.preprocess_bbox(frame, dataset, (res, res)))        # same arithmetic: bit-exact
        ref = z[f"frame_out{i}"]                                                        # cv2's own summation order:
        assert np.all(np.abs(got - ref) <= np.spacing(np.abs(ref).astype(np.float32)))  # <= 1 ulp of float32
        assert np.mean(got != ref) < 1e-3
    # batch of frames, non-square target, down-scaling by a non-integer factor
    g = np.random.RandomState(3)
    frames = g.randint(0, 256, (4, 97, 133, 3)).astype(np.uint8)
    got = ops.preprocess_frames_u8(torch.from_numpy(frames).cuda(), [0.4, 0.5, 0.6], [0.2, 0.25, 0.3], (48, 64)).cpu().numpy()
    for b in range(4):
        x = (frames[b] / 255.0 - np.array([[[0.4, 0.5, 0.6]]])) / np.array([[[0.2, 0.25, 0.3]]])
        want = P.resize_linear_f64(x, (48, 64)).transpose(2, 0, 1).astype(np.float32)
        assert np.array_equal(got[b], want)


def test_estimator_preprocess_bbox_runs_on_device():
    from src.runner.estimator import Estimator
    from oracle.hourglass_oracle import make_state_dict
    cfg = {"MODEL": {"arch": "hg", "num_stacks": 1, "num_classes": 16, "mobile": False, "skip_mode": "sum"},
           "COMMON": {"out_res": 16, "in_res": 64, "dataset": "mpii", "resume": ""}}
    est = Estimator(cfg, state_dict=make_state_dict(num_stacks=1, num_blocks=1, num_classes=16, seed=0))
    frame = np.random.RandomState(5).randint(0, 256, (90, 70, 3)).astype(np.uint8)
    x = est.preprocess_bbox(frame)
    assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == (1, 3, 64, 64)
    assert np.array_equal(x.cpu().numpy(), P.preprocess_bbox(frame, "mpii", (64, 64)))
    kps = est.run(frame)
    assert kps.shape == (16, 2)
    with pytest.raises(TypeError):
        est.preprocess_bbox(frame.astype(np.float32))


def test_u8_pipeline_equals_fp32_pipeline():
    """FlipTestPipeline.infer_host_u8 (uint8 over PCIe, normalise on the device) against infer_host fed with the
    reference's own ToTensor + Normalize output: identical coordinates."""
    from hgb200.infer import FlipTestPipeline
    from src.models import hg
    from oracle.hourglass_oracle import make_state_dict
    sd = make_state_dict(num_stacks=1, num_blocks=1, num_classes=16, seed=0)
    model = hg(num_stacks=1, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum")
    model.load_state_dict(sd)
    model = model.cuda().eval()
    B, H, W = 2, 64, 64
    pipe = FlipTestPipeline(model.engine(), B, H, W)
    pipe.set_affine(np.array([[32.0, 32.0]] * B), np.array([[0.32, 0.32]] * B))
    mean, std = [0.4327, 0.4440, 0.4404], [0.2468, 0.2410, 0.2458]
    g = torch.Generator().manual_seed(9)
    u8 = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(3)]
    f32 = [torch.from_numpy(np.stack([P.normalize_u8(b.numpy(), mean, std) for b in batch])).pin_memory() for batch in u8]
    a = list(pipe.infer_host_u8(u8, mean, std))
    b = list(pipe.infer_host(f32))
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
