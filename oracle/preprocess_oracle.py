"""CPU restatement of the reference's input preprocessing (TEST INFRASTRUCTURE ONLY -- imported by tests/,
__graft_entry__.smoke() and bench.py's CPU legs, never by the package).  SURVEY.md section 8f row N2.

  * normalize_u8         : transforms.ToTensor() + transforms.Normalize(mean, std) on an HWC uint8 crop
                           (src/datasets/common.py:57-64,133-141): float32, x/255 then (x-mean)/std, CHW.
  * resize_linear_f64    : cv2.resize(float64 HWC, dsize, interpolation=INTER_LINEAR) -- the arithmetic lives in
                           OpenCV (unpinned by the reference; 4.13.0 in the dev container), restated from its generic
                           resize: float32 interpolation weights from fx = (float)((d+0.5)*scale-0.5), horizontal
                           pass then vertical pass in float64, no fused multiply-add.
  * preprocess_bbox      : Estimator.preprocess_bbox (src/runner/estimator.py:39-54): /255, per-dataset mean/std
                           in float64 (frame channel order), resize to in_res, CHW, float32.

Pinned against the live reference / OpenCV by oracle/make_golden.py -> tests/golden/preprocess.npz
(tests/test_oracle_golden.py)."""
from __future__ import annotations

import numpy as np

# estimator.py:41-48
DATASET_MEAN_STD = {
    "coco": ([0.4003, 0.4314, 0.4534], [0.2466, 0.2467, 0.2562]),
    "mpii": ([0.4327, 0.4440, 0.4404], [0.2468, 0.2410, 0.2458]),
    "merl": ([0.4785, 0.5036, 0.5078], [0.2306, 0.2289, 0.2326]),
    "se7en11": ([0.5109, 0.5502, 0.5285], [0.2772, 0.2416, 0.2478]),
}


def dataset_mean_std(dataset: str):
    """The first matching branch of the reference's if/elif chain; None when no branch matches (no normalisation)."""
    for key in ("coco", "mpii", "merl", "se7en11"):
        if key in dataset:
            return DATASET_MEAN_STD[key]
    return None


def normalize_u8(img_u8_hwc: np.ndarray, mean, std) -> np.ndarray:
    x = img_u8_hwc.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return ((x - m) / s).astype(np.float32)


def _linear_taps(dst: int, src: int):
    """Source index and float32 weight of the second tap for every destination index (OpenCV resize.cpp)."""
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    idx = np.zeros(dst, dtype=np.int64)
    frac = np.zeros(dst, dtype=np.float32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        idx[d], frac[d] = s, f
    return idx, frac


def resize_linear_f64(img: np.ndarray, dsize) -> np.ndarray:
    """img: float64 [h, w, c]; dsize = (width, height) as cv2 takes it."""
    sh, sw, c = img.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    if (dw, dh) == (sw, sh):
        return img.copy()
    sx, fx = _linear_taps(dw, sw)
    sy, fy = _linear_taps(dh, sh)
    # horizontal: borders use one sample (sx < 0 -> sample 0, weight 0; sx >= w-1 -> last sample alone)
    rows = np.zeros((sh, dw, c), dtype=np.float64)
    for d in range(dw):
        s, f = int(sx[d]), fx[d]
        if s < 0:
            s, f = 0, np.float32(0.0)
        if s >= sw - 1:
            rows[:, d] = img[:, sw - 1] * np.float64(1.0)
            continue
        a0, a1 = np.float64(np.float32(1.0) - f), np.float64(f)
        rows[:, d] = img[:, s] * a0 + img[:, s + 1] * a1
    # vertical: both row indices are clamped, the weights are not changed
    out = np.zeros((dh, dw, c), dtype=np.float64)
    for d in range(dh):
        s, f = int(sy[d]), fy[d]
        s0 = min(max(s, 0), sh - 1)
        s1 = min(max(s + 1, 0), sh - 1)
        b0, b1 = np.float64(np.float32(1.0) - f), np.float64(f)
        out[d] = rows[s0] * b0 + rows[s1] * b1
    return out


def preprocess_bbox(bbox_u8_hwc: np.ndarray, dataset: str, input_size) -> np.ndarray:
    """-> float32 [1, 3, in_res, in_res] exactly as Estimator.preprocess_bbox returns (before .to(device))."""
    x = bbox_u8_hwc / 255.0
    ms = dataset_mean_std(dataset)
    if ms is not None:
        x = (x - np.array([[ms[0]]])) / np.array([[ms[1]]])
    x = resize_linear_f64(x, input_size)
    x = x.transpose((2, 0, 1))
    x = x.reshape((1, 3, input_size[0], input_size[1]))
    return x.astype(np.float32)
