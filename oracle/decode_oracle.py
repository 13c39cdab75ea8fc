"""numpy restatement of the reference heat-map decode, PCK and affine helpers.

Test infrastructure only (see oracle/__init__.py).

Reference followed:
  * get_preds                 src/utils/evaluation.py:8-27
  * calc_dists/dist_acc/accuracy  src/utils/evaluation.py:30-76
  * get_final_preds_v1        src/utils/inference.py:48-67
  * transform_preds / get_affine_transform(inv=1) / affine_transform
                              src/utils/transforms.py:32-94
  * fliplr_joints             src/utils/transforms.py:15-29
  * flip-test averaging       NOT in the reference: defined here
                              from the reference's flip_pairs
                              (src/datasets/mpii.py:29, src/datasets/mscoco.py:59-60)
"""
from __future__ import annotations

import math
import numpy as np

MPII_FLIP_PAIRS = [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]]       # mpii.py:29
COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]  # mscoco.py:59-60


def get_preds(hm: np.ndarray) -> np.ndarray:
    """[B,J,H,W] -> float32 [B,J,2] with the reference's 1-based-index quirk.

    evaluation.py:14-15 : flat argmax, FIRST maximum wins (torch.max / np.argmax).
    evaluation.py:22-23 : x = (idx-1) % W + 1 ; y = floor((idx-1)/W) + 1, float32.
    evaluation.py:25-26 : zeroed where maxval <= 0.
    """
    assert hm.ndim == 4, "Score maps should be 4-dim"
    B, J, H, W = hm.shape
    flat = hm.reshape(B, J, -1)
    idx = flat.argmax(axis=2)
    maxval = np.take_along_axis(flat, idx[..., None], axis=2)[..., 0]
    f = idx.astype(np.float32)
    x = np.mod(f - np.float32(1), np.float32(W)) + np.float32(1)       # python-style mod: (-1)%W = W-1
    y = np.floor((f - np.float32(1)) / np.float32(W)) + np.float32(1)
    preds = np.stack([x, y], axis=2).astype(np.float32)
    preds *= (maxval > 0)[..., None].astype(np.float32)
    return preds


def affine_inv_matrix(center, scale, output_size) -> np.ndarray:
    """get_affine_transform(center, scale, 0, output_size, inv=1), transforms.py:40-73.

    The reference builds three float32 point pairs and calls
    cv2.getAffineTransform(dst, src).  For rot=0 the solution is the closed form
        r  = (scale[0]*200) / output_size[0]          (only scale[0] is used, :49-51)
        x' = cx + (x - ow/2) * r ;  y' = cy + (y - oh/2) * r
    evaluated here from the same float32-rounded points the reference feeds to cv2,
    solved in float64 (cv2 solves the 6x6 system in double).
    """
    center = np.asarray(center, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64) if isinstance(scale, (list, tuple, np.ndarray)) \
        else np.array([scale, scale], dtype=np.float64)
    src_w = scale[0] * 200.0
    dst_w, dst_h = float(output_size[0]), float(output_size[1])
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + np.array([0.0, src_w * -0.5])
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0.0, dst_w * -0.5], np.float32)

    def third(a, b):                    # transforms.py:82-84
        d = a - b
        return b + np.array([-d[1], d[0]], dtype=np.float32)
    src[2] = third(src[0], src[1])
    dst[2] = third(dst[0], dst[1])
    # solve  [dst 1] @ M^T = src  for the 2x3 matrix M (inv=1 maps dst -> src)
    A = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], axis=1)
    M = np.linalg.solve(A, src.astype(np.float64)).T
    return M


def transform_preds(coords: np.ndarray, center, scale, output_size) -> np.ndarray:
    """transforms.py:32-37: per joint t @ [x, y, 1] in float64."""
    M = affine_inv_matrix(center, scale, output_size)
    c = np.asarray(coords, dtype=np.float64)
    out = np.zeros(c.shape, dtype=np.float64)
    out[:, 0] = M[0, 0] * c[:, 0] + M[0, 1] * c[:, 1] + M[0, 2]
    out[:, 1] = M[1, 0] * c[:, 0] + M[1, 1] * c[:, 1] + M[1, 2]
    return out


def quarter_refine(hm_b: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """inference.py:54-61 for one image: hm_b [J,H,W], coords [J,2] (quirk coords)."""
    J, H, W = hm_b.shape
    coords = coords.astype(np.float32).copy()
    for p in range(J):
        hm = hm_b[p]
        px = int(math.floor(coords[p][0] + 0.5))
        py = int(math.floor(coords[p][1] + 0.5))
        if 1 < px < W - 1 and 1 < py < H - 1:
            dx = np.float32(hm[py - 1][px]) - np.float32(hm[py - 1][px - 2])
            dy = np.float32(hm[py][px - 1]) - np.float32(hm[py - 2][px - 1])
            coords[p, 0] += np.float32(np.sign(dx)) * np.float32(0.25)
            coords[p, 1] += np.float32(np.sign(dy)) * np.float32(0.25)
    return coords


def get_final_preds_v1(hms: np.ndarray, center, scale, output_size) -> np.ndarray:
    """inference.py:48-67 -- decodes batch element 0 only, float64 [J,2]."""
    coords = get_preds(hms)[0]
    coords = quarter_refine(hms[0], coords)
    return transform_preds(coords, center, scale, output_size)


def get_final_preds_batch(hms, centers, scales, output_size) -> np.ndarray:
    """The reference routine applied to every batch element independently
    (what a caller looping Estimator.run over a batch gets)."""
    return np.stack([get_final_preds_v1(hms[b:b + 1], centers[b], scales[b], output_size)
                     for b in range(hms.shape[0])])


def accuracy(output: np.ndarray, target: np.ndarray, idxs=None, thr=0.5) -> np.ndarray:
    """PCK on heat maps, evaluation.py:52-76 (incl. the dists[i] quirk at :69)."""
    if idxs is None:
        idxs = list(range(output.shape[1]))
    preds = get_preds(output)
    gts = get_preds(target)
    B, J = preds.shape[:2]
    norm = np.float32(output.shape[3]) / np.float32(10)           # :61, float32 tensor math
    dists = np.zeros((J, B))
    for n in range(B):
        for c in range(J):
            if gts[n, c, 0] > 1 and gts[n, c, 1] > 1:             # :36
                d = preds[n, c].astype(np.float32) - gts[n, c].astype(np.float32)
                dists[c, n] = np.float32(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1]))) / norm
            else:
                dists[c, n] = -1
    acc = np.zeros(len(idxs) + 1)
    avg, cnt = 0.0, 0
    for i in range(len(idxs)):
        d = dists[i]
        d = d[d != -1]
        acc[i + 1] = (1.0 * (d < thr).sum() / len(d)) if len(d) > 0 else -1
        if acc[i + 1] >= 0:
            avg += acc[i + 1]
            cnt += 1
    if cnt != 0:
        acc[0] = avg / cnt
    return acc


def fliplr_joints(joints, joints_vis, width, matched_parts):
    """transforms.py:15-29."""
    joints = joints.copy()
    joints_vis = joints_vis.copy()
    joints[:, 0] = width - joints[:, 0] - 1
    for a, b in matched_parts:
        joints[[a, b]] = joints[[b, a]]
        joints_vis[[a, b]] = joints_vis[[b, a]]
    return joints * joints_vis, joints_vis


def flip_perm(num_joints: int, flip_pairs) -> np.ndarray:
    perm = np.arange(num_joints)
    for a, b in flip_pairs:
        perm[a], perm[b] = b, a
    return perm


def flip_average(hm: np.ndarray, hm_flipped_input: np.ndarray, flip_pairs) -> np.ndarray:
    """SURVEY.md A12 definition: hm_f.flip(-1)[:, perm]; 0.5*(hm + hm_f); no 1-px shift."""
    perm = flip_perm(hm.shape[1], flip_pairs)
    back = hm_flipped_input[:, :, :, ::-1][:, perm]
    return (np.float32(0.5) * (hm.astype(np.float32) + back.astype(np.float32))).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# DARK-style decode: get_final_preds_v2, src/utils/inference.py:9-45,70-87  
# ---------------------------------------------------------------------------------------------
# cv2.getGaussianKernel(11, 0) in float64 (sigma = 0.3*((11-1)*0.5 - 1) + 0.8 = 2.0), OpenCV 4.13.0 -- the arithmetic
# lives in OpenCV (unpinned by the reference); the eleven weights are recorded here so that the restatement does
# not depend on cv2.
GAUSS11 = np.array([float.fromhex(v) for v in (
    '0x1.20c2564ee6772p-7', '0x1.bcb86a082c301p-6', '0x1.0ab50979aaf94p-4', '0x1.f2464c62edaf4p-4',
    '0x1.6a7e1d504a91dp-3', '0x1.9ac20a36ea596p-3', '0x1.6a7e1d504a91dp-3', '0x1.f2464c62edaf4p-4',
    '0x1.0ab50979aaf94p-4', '0x1.bcb86a082c301p-6', '0x1.20c2564ee6772p-7')])


def blur11_zero_padded(hm: np.ndarray) -> np.ndarray:
    """inference.py:37-43 for one map: the map is embedded in a zero frame 5 px wide, blurred with the separable
    11-tap Gaussian in float64 and cropped back, so the frame's own border mode never reaches the crop: a
    zero-padded convolution.  Rows first (taps left to right), then columns (centre, then symmetric pairs); OpenCV's
    exact summation order is not documented -- agreement with cv2.GaussianBlur is ~1e-16, which the reference's own
    store into a float32 array erases except for a rare last-bit flip."""
    h, w = hm.shape
    k = GAUSS11
    a = np.zeros((h, w + 10), dtype=np.float64)
    a[:, 5:-5] = hm
    rows = np.zeros((h + 10, w), dtype=np.float64)
    acc = k[0] * a[:, 0:w]
    for t in range(1, 11):
        acc = acc + k[t] * a[:, t:t + w]
    rows[5:-5] = acc
    out = k[5] * rows[5:5 + h]
    for t in range(1, 6):
        out = out + k[5 + t] * (rows[5 + t:5 + t + h] + rows[5 - t:5 - t + h])
    return out


def gaussian_blur(hm: np.ndarray) -> np.ndarray:
    """inference.py:32-45 on float32 [B,J,h,w]: blur, then rescale so the maximum is the original maximum; every
    store goes through the float32 array (as in the reference, where hms is the float32 view of the tensor)."""
    hm = hm.astype(np.float32).copy()
    B, J = hm.shape[:2]
    for i in range(B):
        for j in range(J):
            origin_max = np.max(hm[i, j])
            hm[i, j] = blur11_zero_padded(hm[i, j].astype(np.float64))
            hm[i, j] *= origin_max / np.max(hm[i, j])
    return hm


def taylor(hm: np.ndarray, coord: np.ndarray) -> np.ndarray:
    """inference.py:9-29; hm float32 [h,w] (log heat map), coord float32 [2] (quirk coordinates)."""
    H, W = hm.shape
    px, py = int(coord[0]), int(coord[1])
    coord = coord.astype(np.float32).copy()
    if 1 < px < W - 2 and 1 < py < H - 2:
        dx = 0.5 * (hm[py][px + 1] - hm[py][px - 1])
        dy = 0.5 * (hm[py + 1][px] - hm[py - 1][px])
        dxx = 0.25 * (hm[py][px + 2] - 2 * hm[py][px] + hm[py][px - 2])
        dxy = 0.25 * (hm[py + 1][px + 1] - hm[py - 1][px + 1] - hm[py + 1][px - 1] + hm[py - 1][px - 1])
        dyy = 0.25 * (hm[py + 2][px] - 2 * hm[py][px] + hm[py - 2][px])
        derivative = np.array([[dx], [dy]])
        hessian = np.array([[dxx, dxy], [dxy, dyy]])
        if dxx * dyy - dxy ** 2 != 0:
            offset = -np.dot(np.linalg.inv(hessian), derivative)
            coord = (coord + np.squeeze(np.array(offset.T), axis=0).astype(np.float32)).astype(np.float32)
    return coord


def get_final_preds_v2(hms: np.ndarray, center, scale, output_size, refine_joints: int = 2) -> np.ndarray:
    """inference.py:70-87 -- batch element 0 only.  The reference's loop is `for p in range(coords.shape[1])` with
    coords of shape [J, 2], so only joints 0 and 1 get the Taylor step (refine_joints=2 reproduces that; pass J for
    what the loop presumably meant)."""
    coords = get_preds(hms)[0]
    h = gaussian_blur(hms[:1])
    h = np.log(np.maximum(h, np.float32(1e-10)))
    for p in range(min(refine_joints, coords.shape[0])):
        coords[p] = taylor(h[0][p], coords[p])
    return transform_preds(coords, center, scale, output_size)
