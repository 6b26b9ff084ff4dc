"""Age-stage crop geometry: the scalar part of ``normalize_image`` (reference ``face_normalization_tools.py:150-311``,
called from ``face_analysis.py:1212-1222`` with ``eyes_inferred-mouth_areaZ`` / ``mid_eyes_inferred-mouth`` /
``EyeLineRotation`` / integer rotation centre / out_size (256, 260)) for a batch of faces, in float64 and in the
reference's operation order, plus the constant NEAREST index tables of the 96 x 96 sub-sampling
(``face_analysis.py:1183-1199, 1230-1246``).  The resampling itself -- integer crop -> BICUBIC rotation -> BICUBIC
EXTENT -> NEAREST sub-sampling -- is ONE kernel (``csrc/crop.cu: age_crop_kernel``) that evaluates, per output sample,
the chain of uint8 images Pillow would materialise (same arithmetic, same roundings) without materialising them.

Un-vendored pieces (cuicuilco ``rotate_improved``, ``load_image_data_monoprocessor``) follow the definitions of
``oracle/normalize.py`` (PARITY UNPINNED there, Pillow-pinned for everything Pillow does)."""
from __future__ import annotations

import math

import numpy as np

OUT_SIZE = (256, 260)
AGE_OBJ_AVG, AGE_OBJ_STD = 0.0, 0.16
FACE_PARAMS = 16       # doubles per face handed to the kernel


def face_params(eyes, im_width, im_height, out_size=OUT_SIZE):
    """eyes: (N, 4) = left x, left y, right x, right y; im_width / im_height: scalars or (N,) arrays.
    Returns (N, 16) float64: crop origin x, y (integers), crop width, height, rotate flag, Pillow's rotation matrix
    a0..a5, and the EXTENT affine xs, x0, ys, y0 of the last resampling; one spare."""
    eyes = np.asarray(eyes, dtype=np.float64).reshape(-1, 4)
    n = len(eyes)
    W = np.broadcast_to(np.asarray(im_width, dtype=np.float64), (n,))
    H = np.broadcast_to(np.asarray(im_height, dtype=np.float64), (n,))
    elx, ely, erx, ery = eyes[:, 0], eyes[:, 1], eyes[:, 2], eyes[:, 3]
    if (elx > erx).any():
        raise ValueError("Warning: the eyes are ordered incorrectly!!!")           # the reference calls exit()
    eyes_x_m = (erx + elx) / 2.0
    eyes_y_m = (ery + ely) / 2.0
    dist_eyes = np.sqrt((elx - erx) ** 2 + (ely - ery) ** 2)
    desired_area = 37.0 * 42.0 / 2.0 * (37.5 / 37.0) ** 2
    eye_dx = erx - elx
    eye_dy = ery - ely
    mouth_x = (erx + elx) / 2.0 - (42.0 / 37.0) * eye_dy
    mouth_y = (ery + ely) / 2.0 + (42.0 / 37.0) * eye_dx
    height = np.sqrt((eyes_x_m - mouth_x) ** 2 + (eyes_y_m - mouth_y) ** 2)
    area = dist_eyes * height / 2.0
    mid_x = (eyes_x_m + mouth_x) / 2.0
    mid_y = (eyes_y_m + mouth_y) / 2.0
    scale_factor = np.sqrt(area / desired_area)
    ori_width = out_size[0] * scale_factor / 2
    ori_height = out_size[1] * scale_factor / 2
    cx_int = np.trunc(mid_x + 0.5)                      # int(): truncation toward zero
    cy_int = np.trunc(mid_y + 0.5)
    win_w = 2 * np.maximum(W - 1 - cx_int + 0.5, cx_int + 0.5)
    win_h = 2 * np.maximum(H - 1 - cy_int + 0.5, cy_int + 0.5)
    Delta_x = mid_x - cx_int
    Delta_y = mid_y - cy_int
    out = np.zeros((n, FACE_PARAMS))
    for k in range(n):                                  # per face: math.atan2 / round() exactly as Python evaluates them
        angle = math.atan2(ery[k] - ely[k], erx[k] - elx[k]) * 180 / math.pi
        rad = -angle * np.pi / 180.0
        dxr = Delta_x[k] * np.cos(rad) - Delta_y[k] * np.sin(rad)
        dyr = Delta_y[k] * np.cos(rad) + Delta_x[k] * np.sin(rad)
        cw, ch = int(win_w[k] + 0.5), int(win_h[k] + 0.5)
        new_cx = (cw - 1) / 2.0 + dxr
        new_cy = (ch - 1) / 2.0 + dyr
        x0 = new_cx - (ori_width[k] - 1) / 2.0
        x1 = (new_cx + (ori_width[k] - 1) / 2.0) + 1
        y0 = new_cy - (ori_height[k] - 1) / 2.0
        y1 = (new_cy + (ori_height[k] - 1) / 2.0) + 1
        out[k, 0] = cx_int[k] - (win_w[k] - 1) / 2.0
        out[k, 1] = cy_int[k] - (win_h[k] - 1) / 2.0
        out[k, 2], out[k, 3] = cw, ch
        a = angle % 360.0
        if a == 0:
            out[k, 4] = 0.0                              # Pillow's rotate(0) is a copy
        elif a == 180:
            out[k, 4] = 2.0                              # transpose(ROTATE_180)
        else:
            out[k, 4] = 1.0
            center = (cw / 2, ch / 2)
            t = -math.radians(a)
            m = [round(math.cos(t), 15), round(math.sin(t), 15), 0.0, round(-math.sin(t), 15), round(math.cos(t), 15), 0.0]
            m[2] = m[0] * (-center[0]) + m[1] * (-center[1]) + m[2]
            m[5] = m[3] * (-center[0]) + m[4] * (-center[1]) + 0.0
            m[2] += center[0]
            m[5] += center[1]
            out[k, 5:11] = m
        out[k, 11] = (x1 - x0) / out_size[0]            # Image.transform: EXTENT -> AFFINE (xs, 0, x0, 0, ys, y0)
        out[k, 12] = x0
        out[k, 13] = (y1 - y0) / out_size[1]
        out[k, 14] = y0
    return out


def age_box(age_subimage_width=96, age_subimage_height=96, out_size=OUT_SIZE):
    """Sampling box of the age patch inside the normalised image (``face_analysis.py:1183-1199``; translations are in
    sampled pixels, ``trans_sampled=True``)."""
    age_image_width, age_image_height = out_size
    reduction_factor = 160.0 / 96
    age_sampling = 1.14 * reduction_factor
    first_row = age_image_height / 2.0 - age_subimage_height * age_sampling / 2.0
    first_column = age_image_width / 2.0 - age_subimage_width * age_sampling / 2.0
    x0 = first_column + (0.0 / reduction_factor) * age_sampling
    y0 = first_row + (-6.0 / reduction_factor) * age_sampling
    return (x0, y0, x0 + age_subimage_width * age_sampling, y0 + age_subimage_height * age_sampling)


def nearest_table(lo, hi, n_out, size):
    """Pillow's NEAREST source index per output sample (sequential double accumulation, -1 = outside)."""
    a = (np.float64(hi) - np.float64(lo)) / np.float64(n_out)
    xo = np.float64(lo) + a * np.float64(0.5)
    idx = np.empty(n_out, dtype=np.int32)
    for c in range(n_out):
        idx[c] = int(xo) if (not xo < 0.0) and xo < size else -1
        xo = xo + a
    return idx


def age_tables(age_subimage_width=96, age_subimage_height=96, out_size=OUT_SIZE):
    x0, y0, x1, y1 = age_box(age_subimage_width, age_subimage_height, out_size)
    return nearest_table(x0, x1, age_subimage_width, out_size[0]), nearest_table(y0, y1, age_subimage_height, out_size[1])
