"""Generate tests/golden/preprocess.npz from the LIVE reference (dev container only; needs /root/reference, cv2 and
torchvision):  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_preprocess.py

  * Estimator.preprocess_bbox (src/runner/estimator.py:39-54) called unbound on seeded uint8 frames of several
    sizes (up-scaling, down-scaling, non-square, identity) for each dataset branch;
  * JointsDataset._get_transformation(mean, std) = ToTensor + Normalize (src/datasets/common.py:57-64) on uint8 crops.
SURVEY.md section 8f row N2."""
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
for name in ("pycocotools", "pycocotools.coco", "torchsummary", "progress", "progress.bar"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["pycocotools.coco"].COCO = object
sys.modules["torchsummary"].summary = lambda *a, **k: None
sys.modules["progress.bar"].Bar = object

from src.runner.estimator import Estimator as RefEstimator        # noqa: E402
from src.datasets.common import JointsDataset                      # noqa: E402

FRAMES = [("mpii", 37, 53, 64), ("coco", 120, 90, 64), ("mpii", 64, 64, 64), ("merl", 200, 300, 128),
          ("se7en11", 5, 7, 64), ("other", 90, 70, 64)]
CROPS = [(64, 64), (128, 64)]


def main():
    rng = np.random.RandomState(11)
    out = {"n_frames": np.array(len(FRAMES)), "n_crops": np.array(len(CROPS))}
    for i, (dataset, fh, fw, res) in enumerate(FRAMES):
        frame = rng.randint(0, 256, (fh, fw, 3)).astype(np.uint8)
        me = types.SimpleNamespace(dataset=dataset, input_size=(res, res), device=torch.device("cpu"))
        y = RefEstimator.preprocess_bbox(me, frame)
        out[f"frame{i}"], out[f"frame_out{i}"] = frame, y.numpy()
        out[f"frame_cfg{i}"] = np.array([dataset, str(res)])
    mean, std = torch.tensor([0.4327, 0.4440, 0.4404]), torch.tensor([0.2468, 0.2410, 0.2458])
    tf = JointsDataset._get_transformation(mean, std)
    for i, (h, w) in enumerate(CROPS):
        crop = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        out[f"crop{i}"], out[f"crop_out{i}"] = crop, tf(crop).numpy()
    out["crop_mean"], out["crop_std"] = mean.numpy(), std.numpy()
    path = os.path.join(REPO, "tests", "golden", "preprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
