"""Final-prediction decode with the reference's entry point (src/utils/inference.py:48-67): arg-max,
quarter-pixel sign shift and inverse affine in ONE kernel launch (hg_decode_final_preds).
`get_final_preds_v2` (inference.py:70-87, DARK-style: Gaussian blur, log, Taylor step) is hg_decode_final_preds_v2."""
import numpy as np

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
    """Reference semantics (inference.py:70-87): batch element 0 only, and -- because the reference's loop runs over
    coords.shape[1] == 2 -- only joints 0 and 1 receive the Taylor refinement; float64 ndarray [J,2].  Unlike the
    reference this does not blur the caller's tensor in place (its `hms.numpy()` aliases the argument)."""
    out = ops.decode_final_preds_v2(hms[0:1], np.asarray(center, dtype=np.float64).reshape(1, 2),
                                    np.asarray(scale, dtype=np.float64).reshape(1, 2), output_size, refine_joints=2)
    return out[0].cpu().numpy()


def get_final_preds_v2_batch(hms, centers, scales, output_size, refine_joints=None):
    """DARK decode for every image and (by default) every joint: float64 ndarray [B,J,2]."""
    rj = hms.shape[1] if refine_joints is None else refine_joints
    return ops.decode_final_preds_v2(hms, centers, scales, output_size, refine_joints=rj).cpu().numpy()
