"""numpy restatement of the JointsMSE loss and the Gaussian target generator.

Test infrastructure only (see oracle/__init__.py).

Reference followed:
  * MSELoss._compute / forward          src/loss/mse.py:19-44
  * JointsDataset.generate_target       src/datasets/common.py:197-248
"""
from __future__ import annotations

import numpy as np


def joints_mse(outputs, target, target_weight, use_target_weight=True):
    """Sum over stacks of  (1/J) * sum_j 0.5 * mean_{b,hw}((w*p - w*g)^2)   (mse.py:27-44).

    outputs: list of [B,J,H,W]; target [B,J,H,W]; target_weight [B,J,1].
    Accumulated in float64 here; the reference accumulates float32 -- compare with rtol.
    Returns (loss, [grad per stack]) with grad = dL/d outputs[s].
    """
    B, J = target.shape[:2]
    hw = target.shape[2] * target.shape[3]
    w = target_weight.reshape(B, J, 1, 1).astype(np.float64) if use_target_weight \
        else np.ones((B, J, 1, 1))
    loss = 0.0
    grads = []
    for o in outputs:
        d = (o.astype(np.float64) - target.astype(np.float64)) * w
        loss += 0.5 * (d * d).sum() / (B * hw) / J
        grads.append((d * w / (B * hw * J)).astype(np.float32))
    return loss, grads


def generate_target(joints, joints_vis, image_size=(256, 256), heatmap_size=(64, 64), sigma=1):
    """common.py:197-248.  joints [J,3] in input-pixel coords, joints_vis [J,3].

    image_size / heatmap_size are (w, h) as in the reference (common.py:45-47).
    Returns target [J,h,w] float32, target_weight [J,1] float32.
    """
    J = joints.shape[0]
    image_size = np.asarray(image_size)
    heatmap_size = np.asarray(heatmap_size)
    target_weight = np.ones((J, 1), dtype=np.float32)
    target_weight[:, 0] = joints_vis[:, 0]
    target = np.zeros((J, heatmap_size[1], heatmap_size[0]), dtype=np.float32)
    tmp = sigma * 3
    size = 2 * tmp + 1
    ax = np.arange(0, size, 1, np.float32)
    g = np.exp(-((ax[None, :] - size // 2) ** 2 + (ax[:, None] - size // 2) ** 2) / (2 * sigma ** 2))
    for j in range(J):
        stride = image_size / heatmap_size
        mu_x = int(joints[j][0] / stride[0] + 0.5)      # int() truncates toward zero (:218-219)
        mu_y = int(joints[j][1] / stride[1] + 0.5)
        ul = [int(mu_x - tmp), int(mu_y - tmp)]
        br = [int(mu_x + tmp + 1), int(mu_y + tmp + 1)]
        if ul[0] >= heatmap_size[0] or ul[1] >= heatmap_size[1] or br[0] < 0 or br[1] < 0:
            target_weight[j] = 0
            continue
        gx = max(0, -ul[0]), min(br[0], heatmap_size[0]) - ul[0]
        gy = max(0, -ul[1]), min(br[1], heatmap_size[1]) - ul[1]
        ix = max(0, ul[0]), min(br[0], heatmap_size[0])
        iy = max(0, ul[1]), min(br[1], heatmap_size[1])
        if target_weight[j] > 0.5:
            target[j][iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    return target, target_weight


def generate_target_batch(joints, joints_vis, image_size=(256, 256), heatmap_size=(64, 64), sigma=1):
    ts, ws = [], []
    for b in range(joints.shape[0]):
        t, w = generate_target(joints[b], joints_vis[b], image_size, heatmap_size, sigma)
        ts.append(t)
        ws.append(w)
    return np.stack(ts), np.stack(ws)
