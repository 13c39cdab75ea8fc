"""Heat-map decode and PCK with the reference's entry points (src/utils/evaluation.py:8-91),
computed by libhgb200's warp-shuffle arg-max kernels.  Results are bit-identical to the reference's
(integer-valued coordinates incl. the 1-based index quirk of evaluation.py:22-23)."""
from __future__ import print_function

import numpy as np
import torch

from hgb200 import ops

__all__ = ['accuracy', 'AverageMeter']


def get_preds(batch_heatmaps):
    """ Input: batch_heatmaps in torch Tensor [batch, njoint, height, width]
        Output: coords of joint [batch, njoint, 2]   (float32, on the heat maps' device)
    """
    assert batch_heatmaps.dim() == 4, 'Score maps should be 4-dim'
    if batch_heatmaps.shape[0] == 0 or batch_heatmaps.shape[1] == 0:     # torch.max over an empty batch: empty result
        return torch.zeros(batch_heatmaps.shape[0], batch_heatmaps.shape[1], 2, dtype=torch.float32,
                           device=batch_heatmaps.device)
    preds, _, _ = ops.decode_argmax(batch_heatmaps)
    return preds if batch_heatmaps.is_cuda else preds.cpu()


def calc_dists(preds, target, normalize):
    """Kept for API compatibility (reference evaluation.py:30-40); `accuracy` no longer calls it --
    the per-(b,j) distances come from one kernel launch instead of a Python double loop."""
    preds = preds.float().cpu()
    target = target.float().cpu()
    dists = np.zeros((preds.size(1), preds.size(0)))
    for n in range(preds.size(0)):
        for c in range(preds.size(1)):
            if target[n, c, 0] > 1 and target[n, c, 1] > 1:
                dists[c, n] = torch.dist(preds[n, c, :], target[n, c, :]) / normalize[n]
            else:
                dists[c, n] = -1
    return dists


def dist_acc(dists, thr=0.5):
    """ Return percentage below threshold while ignoring values with a -1 """
    dist = dists[dists != -1]
    if len(dist) > 0:
        return 1.0 * (dist < thr).sum().item() / len(dist)
    else:
        return -1


def accuracy(output, target, idxs=None, thr=0.5):
    """
    Calculate accuracy according to PCK, but uses ground truth heatmap rather than x,y locations
    First value to be returned is average accuracy across 'idxs', followed by individual accuracies
    (reference evaluation.py:52-76, including its `dists[i]` indexing).
    """
    if idxs is None:
        idxs = list(range(output.shape[1]))
    dists = ops.pck_dists(output, target).cpu().numpy().astype(np.float64)     # [J, B]; one D2H copy

    acc = np.zeros((len(idxs) + 1))
    avg_acc = 0
    cnt = 0

    for i in range(len(idxs)):
        acc[i + 1] = dist_acc(dists[i], thr=thr)
        if acc[i + 1] >= 0:
            avg_acc = avg_acc + acc[i + 1]
            cnt += 1

    if cnt != 0:
        acc[0] = avg_acc / cnt
    return acc


class AverageMeter(object):
    """Computes and stores the average and current value"""
    def __init__(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
