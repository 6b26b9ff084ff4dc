"""Sliding-window grid (host side of the hot path).

Replaces ``compute_sampling_values`` / ``compute_posX_posY_values`` /
``compute_subimage_coordinates_from_posX_posY_values`` (reference ``face_analysis.py:575-669``; SURVEY.md rows
a-1..a-3).  The arithmetic is the reference's float64 operation order -- window enumeration must be
bit-exact because it drives every gather index -- but the Python double loop over windows
(``face_analysis.py:665-668``, the author's own TODO) is a broadcast, and all scales of an image are
returned as one array so that the whole pyramid travels through the cascade as one batch.
"""
from __future__ import annotations

import numpy as np


def compute_sampling_values(im_width, im_height, subimage_width, subimage_height, smallest_face, net_mins, net_maxs,
                            patch_overlap_sampling=1.1, adaptive_grid_scale=True):
    min_side = min(im_height, im_width)
    min_box_side = max(20, min_side * smallest_face * 0.825 / net_mins)
    sampling_value = min_box_side * 1.0 / subimage_width
    if not adaptive_grid_scale:
        return [sampling_value]
    step = (net_maxs / net_mins) / patch_overlap_sampling
    values = []
    while (subimage_width * sampling_value * net_mins / 0.825 < im_width) and (
            subimage_height * sampling_value * net_mins / 0.825 < im_height):
        values.append(sampling_value)
        sampling_value *= step
    return values


def compute_posX_posY_values(im_width, im_height, subimage_width, subimage_height, regression_width,
                             regression_height, sampling_value, net_Dx, net_Dy, patch_overlap_posx_posy=1.1):
    patch_width = subimage_width * sampling_value
    patch_height = subimage_height * sampling_value
    sep_x = net_Dx * 2.0 * patch_width / regression_width
    sep_y = net_Dy * 2.0 * patch_height / regression_height
    # numpy.ceil(...) was handed to linspace as `num` (Python 2 / old numpy truncated it to int)
    n_x = int(np.ceil((1 + (im_width - patch_width) / sep_x) * patch_overlap_posx_posy))
    n_y = int(np.ceil((1 + (im_height - patch_height) / sep_y) * patch_overlap_posx_posy))
    posX = np.linspace(0.0, im_width - patch_width, n_x)
    posY = np.linspace(0.0, im_height - patch_height, n_y)
    return posX, posY, patch_width, patch_height, net_Dx * patch_width / regression_width, \
        net_Dy * patch_height / regression_height


def subimage_coordinates(posX, posY, patch_width, patch_height):
    """Row-major grid, y outer / x inner: row j*n_x+i = [posX_i, posY_j, posX_i+pw-1, posY_j+ph-1]."""
    X = np.broadcast_to(posX[None, :], (len(posY), len(posX))).reshape(-1)
    Y = np.broadcast_to(posY[:, None], (len(posY), len(posX))).reshape(-1)
    coords = np.empty((X.shape[0], 4))
    coords[:, 0] = X
    coords[:, 1] = Y
    coords[:, 2] = X + patch_width - 1
    coords[:, 3] = Y + patch_height - 1
    return coords


def prescaled_size(width, height, prescale_size=1000):
    """``FaceDetectUpdated.py:551-559``: images are first shrunk so that the longer side is <= 1000 px."""
    f = max(width * 1.0 / prescale_size, height * 1.0 / prescale_size)
    if f > 1.0:
        return int(width / f), int(height / f), f
    return width, height, 1.0


def window_pyramid(im_width, im_height, header, smallest_face, overlap_sampling=1.1, overlap_posxy=1.1):
    """All windows of all scales of one image.

    header = (net_Dx, net_Dy, net_Dang, net_mins, net_maxs, subimage_w, subimage_h, regression_w, regression_h)
    Returns dict(coords (N,4) f64, patch_wh (N,2) f64, scale (N,) int32, sampling_values, counts).
    """
    net_Dx, net_Dy, _, net_mins, net_maxs, sw, sh, rw, rh = header
    svals = compute_sampling_values(im_width, im_height, sw, sh, smallest_face, net_mins, net_maxs, overlap_sampling)
    coords, wh, scale, counts = [], [], [], []
    for k, s in enumerate(svals):
        px, py, pw, ph, _, _ = compute_posX_posY_values(im_width, im_height, sw, sh, rw, rh, s, net_Dx, net_Dy, overlap_posxy)
        c = subimage_coordinates(px, py, pw, ph)
        coords.append(c)
        wh.append(np.tile(np.array([[pw, ph]]), (len(c), 1)))
        scale.append(np.full(len(c), k, dtype=np.int32))
        counts.append(len(c))
    if not coords:
        return dict(coords=np.zeros((0, 4)), patch_wh=np.zeros((0, 2)), scale=np.zeros(0, dtype=np.int32),
                    sampling_values=[], counts=[])
    return dict(coords=np.concatenate(coords), patch_wh=np.concatenate(wh), scale=np.concatenate(scale),
                sampling_values=svals, counts=counts)
