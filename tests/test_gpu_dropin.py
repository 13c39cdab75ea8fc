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


# ------------------------------------------------------------------------------------------------ stand-alone modules
def test_standalone_bottleneck_and_hourglass_modules_match_the_oracle():
    """modules.py:27-47 and :80-99 called on their own (eval mode): same kernels as the network plan, fp32 oracle as
    the yardstick (bf16 storage: 2e-2 of the peak)."""
    from src.models.modules import HGBottleneck, Hourglass
    from oracle.hourglass_oracle import bottleneck, hourglass
    torch.manual_seed(3)
    for mobile, skip in ((False, "sum"), (True, "concat")):
        blk = HGBottleneck(256, 128, mobile=mobile).cuda().eval()
        hgm = Hourglass(HGBottleneck, 1, 128, 4, mobile, skip_mode=skip).cuda().eval()
        with torch.no_grad():
            for m in list(blk.modules()) + list(hgm.modules()):
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.normal_(0, 0.1)
                    m.running_var.uniform_(0.5, 1.5)
                    m.weight.uniform_(0.75, 1.25)
                    m.bias.normal_(0, 0.1)
        x = torch.randn(2, 256, 32, 32, generator=torch.Generator().manual_seed(4))
        with torch.no_grad():
            got = blk(x.cuda()).cpu()
            ref = bottleneck({"b." + k: v.cpu() for k, v in blk.state_dict().items()}, "b", x)
            assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
            got = hgm(x.cuda()).cpu()
            ref = hourglass({"m." + k: v.cpu() for k, v in hgm.state_dict().items()}, "m", 4, x)
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max()), (mobile, skip)
    blk.train()
    with pytest.raises(RuntimeError):
        blk(x.cuda())


# ------------------------------------------------------------------------------------------------ the reference's own script
def test_reference_estimate_script_runs_unmodified_over_this_src(tmp_path, monkeypatch, capsys):
    """scripts/estimate.py:7-24 of the reference, byte for byte (vendored into oracle/_ref by oracle/vendor_reference.py),
    executed with THIS repo's `src` package on the path: yaml config -> Estimator(cfg) -> checkpoint in the reference's
    format -> cv2.imread -> estimator.run(frame) -> circles drawn -> cv2.imwrite."""
    import runpy
    import sys
    import cv2
    import yaml
    from oracle.vendor_reference import verify, DEST
    from oracle.hourglass_oracle import make_state_dict
    script = os.path.join(DEST, "scripts", "estimate.py")
    if not (verify() and os.path.isfile(script)):
        pytest.skip("oracle/_ref is not vendored in this checkout (python oracle/vendor_reference.py)")
    sd = make_state_dict(num_stacks=2, num_blocks=1, num_classes=17, mobile=True, skip_mode="sum", seed=5)
    ckpt = tmp_path / "checkpoint.pth.tar"
    torch.save({"epoch": 1, "best_acc": 0.0, "state_dict": {"module." + k: v for k, v in sd.items()}}, str(ckpt))
    frame = np.random.RandomState(0).randint(0, 256, (240, 320, 3), dtype=np.uint8)
    img, dest = tmp_path / "in.png", tmp_path / "out.png"
    cv2.imwrite(str(img), frame)
    cfg = {"MODEL": {"arch": "hg", "num_stacks": 2, "mobile": True, "skip_mode": "sum", "num_classes": 17, "subset": None},
           "COMMON": {"gpu": os.environ.get("CUDA_VISIBLE_DEVICES", "0"), "image_path": str(img), "dest_path": str(dest),
                      "out_res": 64, "in_res": 256, "dataset": "mscoco", "resume": str(ckpt)}}
    cfg_path = tmp_path / "inference.yaml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    monkeypatch.setattr(sys, "argv", [script, str(cfg_path)])
    runpy.run_path(script, run_name="__main__")
    out = capsys.readouterr().out
    assert "creating model 'hg', stacks=2" in out and "Inference time on cuda" in out
    drawn = cv2.imread(str(dest))
    assert drawn is not None and drawn.shape == frame.shape and (drawn != frame).any()      # the key points were drawn


def _write_mpii_fixture(root, n_train=8, n_val=4, seed=0):
    """A tiny dataset in the MPII layout the reference's `mpii` class reads (datasets/mpii.py:42-88): images/*.png,
    annot/{train,valid}.json with 1-based joints, person centre and scale (height / 200)."""
    import json
    import cv2
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "images"))
    os.makedirs(os.path.join(root, "annot"))
    for split, n in (("train", n_train), ("valid", n_val)):
        anno = []
        for i in range(n):
            img = (rng.rand(300, 300, 3) * 40).astype(np.uint8)
            joints = rng.uniform(60, 240, (16, 2))
            for j, (x, y) in enumerate(joints):
                cv2.circle(img, (int(x), int(y)), 6, (int(37 * j) % 256, int(91 * j) % 256, 255 - 13 * j), -1)
            name = f"{split}_{i:03d}.png"
            cv2.imwrite(os.path.join(root, "images", name), img)
            anno.append({"image": name, "center": [150.0, 150.0], "scale": 1.2, "joints": (joints + 1).tolist(),
                         "joints_vis": [1] * 16})
        with open(os.path.join(root, "annot", f"{split}.json"), "w") as f:
            json.dump(anno, f)


def test_reference_train_script_runs_unmodified_over_this_src(tmp_path, monkeypatch, capsys):
    """scripts/train_and_evaluate.py of the reference, byte for byte (vendored into oracle/_ref), with THIS repo's `src` on
    the path: yaml -> `from src import datasets, models` (the reference's own `mpii` dataset class, re-exported by
    src.datasets, reading an MPII-layout fixture through this build's src.utils.transforms) -> Trainer(cfg, n_joints)
    building its DataLoaders -> trainer.train(): two epochs of the sm_100a training step, evaluation, checkpoints in the
    reference's format.  Environment shims only: numpy >= 1.24 dropped `np.float` (mpii.py:53), `torchsummary` (imported
    by the script, used by its evaluate-only branch) is not installed."""
    import runpy
    import sys
    import types
    import yaml
    from oracle.vendor_reference import verify, DEST
    script = os.path.join(DEST, "scripts", "train_and_evaluate.py")
    if not (verify() and os.path.isfile(script) and os.path.isfile(os.path.join(DEST, "src", "datasets", "mpii.py"))):
        pytest.skip("oracle/_ref is not vendored in this checkout (python oracle/vendor_reference.py)")
    data = tmp_path / "mpii"
    _write_mpii_fixture(str(data))
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/mpii")                       # mpii.py:27: './data/mpii/mean.pth.tar', relative to the working directory
    torch.save({"mean": torch.tensor([0.45, 0.45, 0.45]), "std": torch.tensor([0.25, 0.25, 0.25])}, "data/mpii/mean.pth.tar")
    monkeypatch.setattr(np, "float", float, raising=False)
    monkeypatch.setitem(sys.modules, "torchsummary", types.SimpleNamespace(summary=lambda *a, **k: None))
    monkeypatch.setenv("HG_REFERENCE_SRC", os.path.join(DEST, "src"))
    for name in [m for m in sys.modules if m == "src.datasets" or m.startswith("src.datasets.")]:
        monkeypatch.delitem(sys.modules, name)      # re-import src.datasets with the vendored checkout in view
    import src
    if hasattr(src, "datasets"):
        monkeypatch.delattr(src, "datasets")
    ckdir = tmp_path / "ck"
    cfg = {"DATASET": {"name": "mpii", "image_path": str(data / "images"), "annotation_path": str(data / "annot"),
                       "inp_res": 256, "out_res": 64, "flip": True, "sigma": 1, "scale_factor": 0.25, "rot_factor": 30,
                       "label_type": "Gaussian"},
           "MODEL": {"arch": "hg", "num_stacks": 2, "mobile": False, "skip_mode": "sum", "subset": None},
           "COMMON": {"checkpoint_dir": str(ckdir), "snapshot": 1, "resume": "", "evaluate_only": False, "pck": 0.5,
                      "gpu": os.environ.get("CUDA_VISIBLE_DEVICES", "0")},
           "TRAIN": {"num_workers": 0, "epochs": 1, "start_epoch": 0, "train_batch": 4, "val_batch": 4,
                     "learning_rate": 2.5e-4, "schedule": [1], "gamma": 0.1}}
    cfg_path = tmp_path / "train.yaml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    monkeypatch.setattr(sys, "argv", [script, str(cfg_path)])
    runpy.run_path(script, run_name="__main__")
    out = capsys.readouterr().out
    assert "creating model 'hg', stacks=2" in out and "Epoch: 2" in out
    from src import datasets
    assert datasets.REFERENCE_DATASETS == os.path.join(DEST, "src", "datasets") and datasets.mpii.__module__ == "src.datasets.mpii"
    ck = ckdir / "mpii_hg_s2_non-mobile_all" / "ckpts" / "checkpoint_2.pth.tar"
    assert ck.is_file()
    state = torch.load(str(ck), map_location="cpu", weights_only=False)
    assert state["epoch"] == 2 and all(k.startswith("module.") for k in state["state_dict"])
    assert all(torch.isfinite(v).all() for v in state["state_dict"].values() if v.is_floating_point())
