"""Generate tests/golden/*.npz by running the LIVE reference on seeded inputs.

Run in the dev container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures are committed; the GPU box never sees /root/reference.  The oracle
(oracle/*.py) is checked against these fixtures by tests/test_oracle_golden.py, and
the CUDA path is checked against both.  The reference ships no golden vectors of
its own, so these are the pin.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

# stubs for modules the reference imports but this image lacks
for name in ("pycocotools", "pycocotools.coco", "torchsummary", "progress", "progress.bar"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["pycocotools.coco"].COCO = object
sys.modules["torchsummary"].summary = lambda *a, **k: None
sys.modules["progress.bar"].Bar = object

from src.models.hourglass import hg as ref_hg                      # noqa: E402
from src.loss.mse import MSELoss as RefMSELoss                     # noqa: E402
from src.utils.evaluation import get_preds as ref_get_preds, accuracy as ref_accuracy  # noqa: E402
from src.utils.inference import get_final_preds_v1 as ref_final_v1  # noqa: E402
from src.utils.transforms import get_affine_transform as ref_affine, fliplr_joints as ref_fliplr  # noqa: E402
from src.datasets.common import JointsDataset                      # noqa: E402

from oracle.hourglass_oracle import make_state_dict                # noqa: E402
from oracle.golden_inputs import heatmap_cases, accuracy_cases, loss_cases  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def model_case(name, num_stacks, num_classes, B, H, W, seed, mobile=False, skip_mode="sum",
               num_blocks=1):
    sd = make_state_dict(num_stacks=num_stacks, num_blocks=num_blocks, num_classes=num_classes,
                         mobile=mobile, skip_mode=skip_mode, seed=seed)
    model = ref_hg(num_stacks=num_stacks, num_blocks=num_blocks, num_classes=num_classes,
                   mobile=mobile, skip_mode=skip_mode)
    model.load_state_dict(sd, strict=True)        # proves key names + shapes match the reference
    model.eval()
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(B, 3, H, W, generator=g)
    with torch.no_grad():
        outs = model(x)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg=np.array([num_stacks, num_classes, B, H, W, seed, int(mobile),
                      int(skip_mode == "concat"), num_blocks]),
        n_keys=np.array(len(sd)),
        **{f"out{i}": o.numpy().astype(np.float32) for i, o in enumerate(outs)})
    print(name, "keys", len(sd), "out", tuple(outs[-1].shape), "absmax", float(outs[-1].abs().max()))


def decode_golden():
    """Outputs only; the inputs come from oracle/golden_inputs.py."""
    from oracle.golden_inputs import decode_args, affine_args
    out = {}
    cases = heatmap_cases()
    centers, scales = decode_args(cases)
    for i, hm in enumerate(cases):
        t = torch.from_numpy(hm)
        out[f"preds{i}"] = ref_get_preds(t).numpy()
        B, J, H, W = hm.shape
        out[f"final{i}"] = np.stack([ref_final_v1(t[b:b + 1].clone(), centers[i][b], scales[i][b], (W, H))
                                     for b in range(B)])
    out["n"] = np.array(len(cases))
    args = affine_args()
    out["affine_mats"] = np.stack([ref_affine(np.array(a[0:2]), np.array(a[2:4]), 0, (int(a[4]), int(a[5])), inv=1)
                                   for a in args])
    for i, (pred, tgt) in enumerate(accuracy_cases()):
        out[f"acc{i}"] = ref_accuracy(torch.from_numpy(pred), torch.from_numpy(tgt), None, 0.5)
        out[f"acc_sub{i}"] = ref_accuracy(torch.from_numpy(pred[:, [1, 3, 5]]), torch.from_numpy(tgt[:, [1, 3, 5]]),
                                          [1, 3, 5], 0.5)
        out[f"acc_thr{i}"] = ref_accuracy(torch.from_numpy(pred), torch.from_numpy(tgt), None, 0.2)
    rng = np.random.RandomState(29)
    j = rng.uniform(0, 255, (16, 3))
    v = (rng.rand(16, 1) > 0.3).astype(np.float64).repeat(3, 1)
    fj, fv = ref_fliplr(j.copy(), v.copy(), 256, [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]])
    out["flip_j"], out["flip_v"], out["flip_j_out"], out["flip_v_out"] = j, v, fj, fv
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **out)
    print("decode cases", len(cases))


class _DS(JointsDataset):
    """Bare JointsDataset carrying only what generate_target reads (common.py:197-248)."""
    def __init__(self, J, image_size, heatmap_size, sigma):
        self.num_joints = J
        self.target_type = "Gaussian"
        self.image_size = np.array(image_size)
        self.heatmap_size = np.array(heatmap_size)
        self.sigma = sigma


GRAD_STRIDE = 61   # gradients are stored as a strided sample to keep the fixture small


def loss_golden():
    out = {}
    cases = loss_cases()
    for i, c in enumerate(cases):
        B, J = c["joints"].shape[:2]
        ds = _DS(J, c["isz"], c["hsz"], 1)
        tg, tw = [], []
        for b in range(B):
            t, w = ds.generate_target(c["joints"][b], c["vis"][b])
            tg.append(t)
            tw.append(w)
        tg, tw = np.stack(tg), np.stack(tw)
        outs = [torch.from_numpy(tg + n).requires_grad_(True) for n in c["noise"]]
        loss = RefMSELoss(use_target_weight=True)(outs, torch.from_numpy(tg), torch.from_numpy(tw))
        loss.backward()
        loss_nw = RefMSELoss(use_target_weight=False)([o.detach() for o in outs], torch.from_numpy(tg),
                                                      torch.from_numpy(tw))
        out[f"target{i}"], out[f"tw{i}"] = tg, tw            # sparse -> compresses to a few KB
        for s in range(c["S"]):
            out[f"grad{i}_{s}"] = outs[s].grad.numpy().reshape(-1)[::GRAD_STRIDE].copy()
        out[f"loss{i}"] = np.array(float(loss.detach()))
        out[f"loss_nw{i}"] = np.array(float(loss_nw))
    out["n"] = np.array(len(cases))
    out["grad_stride"] = np.array(GRAD_STRIDE)
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print("loss cases", len(cases))


from oracle.make_golden_inputs import train_inputs, TRAIN_STRIDE  # noqa: E402


def train_golden(name="train_s2_j16_128", num_stacks=2, J=16, B=4, H=128, W=128, seed=7, steps=2, lr=2.5e-4):
    """Two steps of the reference's training loop body (trainer.py:89-99): model.train(), MSELoss(True),
    loss.backward(), RMSprop(lr).step().  Stores losses, strided samples of the last step's gradients, of the
    updated parameters and of the BN running statistics."""
    sd = make_state_dict(num_stacks=num_stacks, num_blocks=1, num_classes=J, seed=seed)
    model = ref_hg(num_stacks=num_stacks, num_blocks=1, num_classes=J, mobile=False, skip_mode="sum")
    model.load_state_dict(sd, strict=True)
    model.train()
    crit = RefMSELoss(use_target_weight=True)
    opt = torch.optim.RMSprop(model.parameters(), lr=lr, momentum=0, weight_decay=0)
    losses = []
    for x, tg, tw in train_inputs(seed + 1, B, J, H, W, steps):
        outs = model(x)
        loss = crit(outs, tg, tw)
        opt.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        opt.step()
        losses.append(float(loss.detach()))
    out = {"cfg": np.array([num_stacks, J, B, H, W, seed, steps]), "lr": np.array(lr), "losses": np.array(losses),
           "stride": np.array(TRAIN_STRIDE)}
    keys = sorted(grads)
    out["grad_sample"] = np.concatenate([grads[k].reshape(-1).numpy() for k in keys])[::TRAIN_STRIDE].copy()
    out["grad_norms"] = np.array([float(grads[k].norm()) for k in keys])
    fsd = model.state_dict()
    out["param_sample"] = np.concatenate([fsd[k].reshape(-1).float().numpy() for k in keys])[::TRAIN_STRIDE].copy()
    bkeys = sorted(k for k in fsd if k.endswith("running_mean") or k.endswith("running_var"))
    out["stat_sample"] = np.concatenate([fsd[k].reshape(-1).numpy() for k in bkeys])[::7].copy()
    out["heat_last"] = outs[-1].detach().numpy()[:, :, ::4, ::4].copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "losses", losses, "params", len(keys))


if __name__ == "__main__":
    torch.set_num_threads(8)
    # C1-shaped (2-stack, 256x256, B=1) plus small/variant shapes
    model_case("model_c1_s2_j16_256", 2, 16, 1, 256, 256, seed=0)
    model_case("model_s2_j17_128x192", 2, 17, 2, 128, 192, seed=1)
    model_case("model_s1_j14_64", 1, 14, 2, 64, 64, seed=2)
    model_case("model_s2_j16_64_mobile", 2, 16, 1, 64, 64, seed=3, mobile=True)
    model_case("model_s2_j16_64_concat", 2, 16, 1, 64, 64, seed=4, skip_mode="concat")
    model_case("model_s1_j16_64_nb2", 1, 16, 1, 64, 64, seed=5, num_blocks=2)
    decode_golden()
    loss_golden()
    train_golden()
