"""Flip-test helpers.  The reference has no flip test; the pairing tables are the
reference's own (src/datasets/mpii.py:29, src/datasets/mscoco.py:59-60)."""
import torch

MPII_FLIP_PAIRS = [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]]
COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]

_cache = {}


def flip_perm_tensor(num_joints: int, flip_pairs, device) -> torch.Tensor:
    key = (num_joints, tuple(map(tuple, flip_pairs)), str(device))
    t = _cache.get(key)
    if t is None:
        perm = list(range(num_joints))
        for a, b in flip_pairs:
            perm[a], perm[b] = b, a
        t = torch.tensor(perm, dtype=torch.int32, device=device)
        _cache[key] = t
    return t
