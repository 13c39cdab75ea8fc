"""Deterministic INPUT generators shared by oracle/make_golden.py and the tests.

Test infrastructure only.  Inputs are regenerated from seeds (numpy RandomState
streams are stable across numpy versions) so that tests/golden/*.npz only has
to carry the reference's OUTPUTS.
"""
from __future__ import annotations

import numpy as np


def heatmap_cases(seed=7):
    """Heat maps exercising every get_preds quirk."""
    rng = np.random.RandomState(seed)
    cases = []
    # (a) hand-made quirk table on a 4x5 map: peaks at (0,0) and (3,2)
    hm = np.zeros((1, 2, 4, 5), np.float32)
    hm[0, 0, 0, 0] = 1.0
    hm[0, 1, 2, 3] = 1.0
    cases.append(hm)
    # (b) random positive maps, several shapes incl. non-square 64x48
    for (B, J, H, W) in [(4, 16, 64, 64), (2, 17, 64, 48), (3, 21, 16, 16), (1, 14, 8, 8), (2, 5, 4, 4)]:
        cases.append(rng.rand(B, J, H, W).astype(np.float32))
    # (c) signed maps (some all-negative), ties, x=0 column peaks, idx=0 peaks
    hm = rng.randn(4, 16, 64, 64).astype(np.float32)
    hm[0, 0] = -np.abs(hm[0, 0]) - 0.1          # all negative -> (0,0)
    hm[0, 1] = 0.0                              # all zero -> maxval 0 -> masked
    hm[0, 2] = 0.0
    hm[0, 2, 10, 20] = 3.0
    hm[0, 2, 30, 5] = 3.0                       # tie: first flat index wins
    hm[0, 3, 17, 0] = 50.0                      # x0 = 0 column
    hm[0, 4, 0, 0] = 50.0                       # idx = 0
    hm[0, 5, 63, 63] = 50.0                     # last element
    hm[0, 6, 0, 63] = 50.0
    hm[0, 7, 63, 0] = 50.0
    hm[1, 0, 1, 2] = 50.0                       # refine guard edges
    hm[1, 1, 61, 62] = 50.0
    hm[1, 2, 62, 61] = 50.0
    hm[1, 3, 5, 1] = 50.0
    hm[1, 4, 5, 2] = 50.0
    cases.append(hm)
    # (d) smooth gaussian blobs (realistic), sub-pixel refine fires in all directions
    ys, xs = np.mgrid[0:64, 0:64].astype(np.float32)
    hm = np.zeros((3, 16, 64, 64), np.float32)
    for b in range(3):
        for j in range(16):
            cx, cy = rng.uniform(-2, 66), rng.uniform(-2, 66)
            hm[b, j] = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 2.0) + 0.01 * rng.randn(64, 64)
    cases.append(hm.astype(np.float32))
    return cases




def decode_args(cases, seed=11):
    """Per-image (center, scale) for get_final_preds_v1, one RandomState stream."""
    rng = np.random.RandomState(seed)
    centers, scales = [], []
    for hm in cases:
        B = hm.shape[0]
        c = np.stack([np.array([rng.uniform(50, 200), rng.uniform(50, 200)]) for _ in range(B)])
        s = np.stack([np.array([rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.5)]) for _ in range(B)])
        centers.append(c)
        scales.append(s)
    return centers, scales


def affine_args(seed=13):
    rng = np.random.RandomState(seed)
    args = []
    for _ in range(8):
        c = [rng.uniform(0, 300), rng.uniform(0, 300)]
        s = [rng.uniform(0.3, 3), rng.uniform(0.3, 3)]
        osz = [(64, 64), (64, 48), (48, 64), (256, 192)][rng.randint(4)]
        args.append([c[0], c[1], s[0], s[1], osz[0], osz[1]])
    args.append([100., 50., 1.5, 2.5, 64, 48])      # SURVEY.md A10 probe
    return np.array(args)


def accuracy_cases(seed=17):
    """(pred, target) heat-map pairs for the PCK routine: prediction = target peak + jitter."""
    rng = np.random.RandomState(seed)
    cases = []
    for (B, J, H, W) in [(8, 16, 64, 64), (4, 17, 64, 48)]:
        tgt = np.zeros((B, J, H, W), np.float32)
        pred = (0.05 * rng.randn(B, J, H, W)).astype(np.float32)
        for b in range(B):
            for j in range(J):
                if rng.rand() < 0.85:
                    y, x = rng.randint(0, H), rng.randint(0, W)
                    tgt[b, j, y, x] = 1.0
                    yy = int(np.clip(y + rng.randint(-4, 5), 0, H - 1))
                    xx = int(np.clip(x + rng.randint(-4, 5), 0, W - 1))
                    pred[b, j, yy, xx] += 1.0
        cases.append((pred, tgt))
    return cases


LOSS_CFGS = [(4, 16, (256, 256), (64, 64), 2), (3, 17, (192, 256), (48, 64), 3), (1, 16, (256, 256), (64, 64), 1)]


def loss_cases(seed=23):
    """joints / visibility (with hand-placed edge cases) and per-stack prediction noise.

    Returns a list of dicts: joints [B,J,3], vis [B,J,3], isz (w,h), hsz (w,h), S, noise list.
    The prediction fed to the loss is ``target + noise[s]`` where target comes from
    generate_target on (joints, vis).
    """
    rng = np.random.RandomState(seed)
    cases = []
    for (B, J, isz, hsz, S) in LOSS_CFGS:
        joints = np.zeros((B, J, 3))
        vis = np.zeros((B, J, 3))
        joints[..., 0] = rng.uniform(-40, isz[0] + 40, (B, J))
        joints[..., 1] = rng.uniform(-40, isz[1] + 40, (B, J))
        v = (rng.rand(B, J) < 0.8).astype(np.float64)
        vis[..., 0] = v
        vis[..., 1] = v
        edge = [[-30.0, 100.0],                 # fully off-map -> weight 0
                [0.0, 0.0],
                [isz[0] - 1, isz[1] - 1],
                [-1.9, -1.9],                   # int() truncation toward zero
                [-14.0, 50.0],                  # br == 0 boundary
                [isz[0] + 11.0, 20.0],          # ul == heatmap_size boundary
                [isz[0] + 9.9, 20.0]]
        for k, e in enumerate(edge):
            joints[0, k, :2] = e
            vis[0, k, :2] = 1
        noise = [(0.3 * rng.randn(B, J, hsz[1], hsz[0])).astype(np.float32) for _ in range(S)]
        cases.append(dict(joints=joints, vis=vis, isz=isz, hsz=hsz, S=S, noise=noise))
    return cases


def dark_cases(seed=29):
    """Inputs for get_final_preds_v2 (batch element 0 is what the reference decodes): Gaussian blobs at sub-pixel
    centres (well-conditioned Hessians), blobs on the border (the guard skips the Taylor step), a non-square map,
    and one plain random map (ill-conditioned: the Taylor step can be large)."""
    rng = np.random.RandomState(seed)
    cases = []

    def blobs(J, H, W, lo, hi, noise):
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
        hm = np.zeros((1, J, H, W), np.float32)
        for j in range(J):
            cx, cy = rng.uniform(lo, W - lo), rng.uniform(lo, H - lo)
            s = rng.uniform(1.0, 2.5)
            hm[0, j] = (hi * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * s * s)) + noise * rng.rand(H, W)).astype(np.float32)
        return hm

    for _ in range(6):
        cases.append(dict(hm=blobs(16, 64, 64, 6.0, rng.uniform(0.3, 1.0), 1e-3)))
    cases.append(dict(hm=blobs(17, 64, 48, 6.0, 0.9, 1e-3)))
    cases.append(dict(hm=blobs(16, 64, 64, -1.0, 0.8, 1e-3)))        # some centres on / beyond the border
    cases.append(dict(hm=blobs(4, 16, 16, 3.0, 0.7, 0.0)))
    z = blobs(16, 64, 64, 6.0, 0.9, 0.0)
    z[0, 1] = 0.0                                                      # empty map: maxval 0 -> coords (0,0), log(1e-10)
    cases.append(dict(hm=z))
    cases.append(dict(hm=rng.rand(1, 16, 64, 64).astype(np.float32)))
    for c in cases:
        H, W = c["hm"].shape[2:]
        c["center"] = np.array([rng.uniform(50, 200), rng.uniform(50, 200)])
        c["scale"] = np.array([rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.5)])
        c["output_size"] = (W, H)
    return cases
