"""Final-prediction decode with the reference's entry point (src/utils/inference.py:48-67): arg-max,
quarter-pixel sign shift and inverse affine in ONE kernel launch (hg_decode_final_preds).
`get_final_preds_v2` (DARK-style refine) is a "next" row."""
import numpy as np
import torch

from hgb200 import ops


def get_final_preds_v1(hms, center, scale, output_size):
    """Reference semantics: decodes batch element 0 only and returns float64 ndarray [J,2]."""
    out = ops.decode_final_preds(hms[0:1], np.asarray(center, dtype=np.float64).reshape(1, 2),
                                 np.asarray(scale, dtype=np.float64).reshape(1, 2), output_size)
    return out[0].cpu().numpy()


def get_final_preds_batch(hms, centers, scales, output_size):
    """The same routine for every image of the batch: float64 ndarray [B,J,2]."""
    return ops.decode_final_preds(hms, centers, scales, output_size).cpu().numpy()


def get_final_preds_v2(hms, center, scale, output_size):
    raise NotImplementedError("DARK-style decode is not on the sm_100a path yet (SURVEY.md 8f N3)")
