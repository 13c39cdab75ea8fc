"""Seeded training batches shared by oracle/make_golden.py (live reference) and the tests (oracle / CUDA).
Test infrastructure only (see oracle/__init__.py)."""
import numpy as np
import torch


def train_inputs(seed, B, J, H, W, steps):
    """Seeded (x, target, target_weight) batches shared by make_golden.py and the tests."""
    from oracle.loss_oracle import generate_target_batch
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(steps):
        x = torch.from_numpy(rng.randn(B, 3, H, W).astype(np.float32))
        joints = np.zeros((B, J, 3))
        joints[..., 0] = rng.uniform(0, W, (B, J))
        joints[..., 1] = rng.uniform(0, H, (B, J))
        vis = (rng.rand(B, J, 1) < 0.8).astype(np.float64).repeat(3, 2)
        tg, tw = generate_target_batch(joints, vis, (W, H), (W // 4, H // 4), 1)
        out.append((x, torch.from_numpy(tg), torch.from_numpy(tw)))
    return out


TRAIN_STRIDE = 997
