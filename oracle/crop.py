"""Oracle (test infrastructure): window -> fixed-size patch resampling.

Reference call chain: ``load_network_subimages`` (``face_analysis.py:775-800``) ->
cuicuilco ``extract_subimages_rotate`` + ``images_asarray`` (un-vendored) ->
Pillow ``Image.transform(out_size, Image.EXTENT, box, filter)``.

angle == 0 (every window of stage Disc1, the hot case): restates Pillow's C resampler and is PINNED
against Pillow 12.2 run live (``tests/test_oracle_crop.py``) and the fixtures in ``tests/golden``:

* NEAREST  (``ImagingScaleAffine``): ``a = (x1 - x0) / ow`` in double; ``xo = x0 + a * 0.5``; per output
  column ``xin = (xo < 0) ? OOB : (int) xo`` then ``xo += a`` -- *sequential* double additions, not
  ``x0 + a * (c + 0.5)`` (Pillow matches the accumulate form 72/72 on the
  adversarial boxes where the two forms differ).  Rows likewise.  OOB (``< 0`` or ``>= size``) -> 0.
* BILINEAR (generic affine path): ``xin = a * (c + 0.5) + x0`` (multiply form); reject if outside
  ``[0, size)``; shift by -0.5; floor; lerp with edge-clamped neighbours; mode 'L' truncates.

* BICUBIC (generic affine path, ``bicubic_filter8``): same source coordinate; reject if outside ``[0, size)``;
  shift by -0.5; floor; 4 x 4 neighbourhood starting one pixel up-left, columns clamped to the image, a row
  outside the image repeats the previous row's value (the first row is clamped); cubic
  ``p1 + d (p2 + d (p3 + d p4))`` with ``p1 = v2, p2 = -v1 + v3, p3 = 2 (v1 - v2) + v3 - v4,
  p4 = -v1 + v2 - v3 + v4``; clipped to [0, 255]; mode 'L' truncates.  PINNED against Pillow 12.2 live.

angle != 0: cuicuilco crops a larger region, rotates it with ``rotate_improved(BICUBIC)`` and re-crops;
that source is not available -- PARITY UNPINNED.  The oracle defines the rotated window as ONE affine
NEAREST/BILINEAR resampling about the box centre (same sampling-point convention as above, multiply
form), which is the transformation the reference composes in two resampling steps.
"""
import numpy as np


def nearest_index_table(lo, hi, n_out, size):
    """Source index per output pixel along one axis, -1 where out of bounds (accumulate form)."""
    a = (np.float64(hi) - np.float64(lo)) / np.float64(n_out)
    xo = np.float64(lo) + a * np.float64(0.5)
    idx = np.empty(n_out, dtype=np.int64)
    for c in range(n_out):
        if xo < 0.0:
            idx[c] = -1
        else:
            xi = int(xo)          # C (int) cast: truncation toward zero, xo >= 0 here
            idx[c] = xi if xi < size else -1
        xo = xo + a
    return idx


def extent_nearest(img, box, out_size=(64, 64)):
    """``img`` uint8 (H, W); ``box`` = (x0, y0, x1, y1) float64; returns uint8 (oh, ow)."""
    ow, oh = out_size
    H, W = img.shape
    xi = nearest_index_table(box[0], box[2], ow, W)
    yi = nearest_index_table(box[1], box[3], oh, H)
    out = np.zeros((oh, ow), dtype=img.dtype)
    vx = xi >= 0
    vy = yi >= 0
    if vx.any() and vy.any():
        sub = img[np.clip(yi, 0, H - 1)][:, np.clip(xi, 0, W - 1)]
        out = np.where(vy[:, None] & vx[None, :], sub, 0).astype(img.dtype)
    return out


def _bilinear_sample(img, xin, yin):
    """Pillow's ``bilinear_filter8`` at continuous source coordinates (arrays of equal shape)."""
    H, W = img.shape
    valid = (xin >= 0.0) & (xin < W) & (yin >= 0.0) & (yin < H)
    xs = xin - 0.5
    ys = yin - 0.5
    x = np.floor(xs).astype(np.int64)
    y = np.floor(ys).astype(np.int64)
    dx = xs - x
    dy = ys - y
    x0 = np.clip(x, 0, W - 1)
    x1 = np.clip(x + 1, 0, W - 1)
    y0 = np.clip(y, 0, H - 1)
    y1 = np.clip(y + 1, 0, H - 1)
    f = img.astype(np.float64)
    v1 = f[y0, x0] + (f[y0, x1] - f[y0, x0]) * dx
    v2 = f[y1, x0] + (f[y1, x1] - f[y1, x0]) * dx
    v = v1 + (v2 - v1) * dy
    return np.where(valid, v, 0.0), valid


def extent_bilinear(img, box, out_size=(64, 64)):
    ow, oh = out_size
    ax = (np.float64(box[2]) - np.float64(box[0])) / ow
    ay = (np.float64(box[3]) - np.float64(box[1])) / oh
    xin = ax * (np.arange(ow) + 0.5) + np.float64(box[0])
    yin = ay * (np.arange(oh) + 0.5) + np.float64(box[1])
    X, Y = np.meshgrid(xin, yin)
    v, _ = _bilinear_sample(img, X, Y)
    return v.astype(np.uint8) if img.dtype == np.uint8 else v     # 'L': truncation


def _cubic(v1, v2, v3, v4, d):
    p1 = v2
    p2 = -v1 + v3
    p3 = 2 * (v1 - v2) + v3 - v4
    p4 = -v1 + v2 - v3 + v4
    return p1 + d * (p2 + d * (p3 + d * p4))


def _bicubic_sample(img, xin, yin):
    """Pillow's ``bicubic_filter8`` at continuous source coordinates (arrays of equal shape) -> uint8."""
    H, W = img.shape
    valid = (xin >= 0.0) & (xin < W) & (yin >= 0.0) & (yin < H)
    xs = xin - 0.5
    ys = yin - 0.5
    xf = np.floor(xs)
    yf = np.floor(ys)
    dx = xs - xf
    dy = ys - yf
    x = xf.astype(np.int64) - 1
    y = yf.astype(np.int64) - 1
    f = img.astype(np.float64)
    xc = [np.clip(x + k, 0, W - 1) for k in range(4)]
    rows = []
    for k in range(4):
        yy = np.clip(y, 0, H - 1) if k == 0 else y + k
        inside = np.ones_like(valid) if k == 0 else (yy >= 0) & (yy < H)
        yr = np.clip(yy, 0, H - 1)
        v = _cubic(f[yr, xc[0]], f[yr, xc[1]], f[yr, xc[2]], f[yr, xc[3]], dx)
        rows.append(v if k == 0 else np.where(inside, v, rows[k - 1]))
    r = _cubic(rows[0], rows[1], rows[2], rows[3], dy)
    out = np.where(r <= 0.0, 0.0, np.where(r >= 255.0, 255.0, np.trunc(r)))
    return np.where(valid, out, 0.0).astype(np.uint8)


def extent_bicubic(img, box, out_size=(64, 64)):
    ow, oh = out_size
    ax = (np.float64(box[2]) - np.float64(box[0])) / ow
    ay = (np.float64(box[3]) - np.float64(box[1])) / oh
    xin = ax * (np.arange(ow) + 0.5) + np.float64(box[0])
    yin = ay * (np.arange(oh) + 0.5) + np.float64(box[1])
    X, Y = np.meshgrid(xin, yin)
    return _bicubic_sample(img, X, Y)


def rotated_sample_points(box, angle_deg, out_size=(64, 64)):
    """Continuous source coordinates of every output pixel for a window rotated by ``angle_deg``
    (the ``delta_ang`` handed to ``extract_subimages_rotate``, i.e. ``-curr_angle``) about its centre."""
    ow, oh = out_size
    x0, y0, x1, y1 = [np.float64(v) for v in box]
    ax = (x1 - x0) / ow
    ay = (y1 - y0) / oh
    cx = (x0 + x1) * 0.5
    cy = (y0 + y1) * 0.5
    u = ax * (np.arange(ow) + 0.5) + x0 - cx
    v = ay * (np.arange(oh) + 0.5) + y0 - cy
    U, V = np.meshgrid(u, v)
    th = np.float64(angle_deg) * np.pi / 180.0
    c, s = np.cos(th), np.sin(th)
    return cx + (U * c - V * s), cy + (U * s + V * c)


def extent_rotated(img, box, angle_deg, out_size=(64, 64), bilinear=False, bicubic=False):
    X, Y = rotated_sample_points(box, angle_deg, out_size)
    H, W = img.shape
    if bicubic:
        return _bicubic_sample(img, X, Y)
    if bilinear:
        v, _ = _bilinear_sample(img, X, Y)
        return v.astype(np.uint8) if img.dtype == np.uint8 else v
    valid = (X >= 0.0) & (Y >= 0.0)
    xi = np.where(valid, X, 0.0).astype(np.int64)
    yi = np.where(valid, Y, 0.0).astype(np.int64)
    valid &= (xi < W) & (yi < H)
    out = np.where(valid, img[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)], 0)
    return out.astype(img.dtype)


NEAREST, BILINEAR, BICUBIC = 0, 2, 3   # Pillow's Image.NEAREST / Image.BILINEAR / Image.BICUBIC values


def extract_subimages(img, coords, angles=None, out_size=(64, 64), interpolation=NEAREST):
    """``load_network_subimages`` result: ``(N, ow*oh)`` float64 with values 0..255 (row-major patches).

    ``angles`` are the *current* face angles; like the reference (``face_analysis.py:781``) the patch is
    extracted with ``delta_ang = -angle``.
    """
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 4)
    n = coords.shape[0]
    ow, oh = out_size
    out = np.zeros((n, ow * oh), dtype=np.float64)
    for k in range(n):
        ang = 0.0 if angles is None else float(angles[k])
        if ang == 0.0:
            p = extent_bicubic(img, coords[k], out_size) if interpolation == BICUBIC \
                else extent_bilinear(img, coords[k], out_size) if interpolation == BILINEAR \
                else extent_nearest(img, coords[k], out_size)
        else:
            p = extent_rotated(img, coords[k], -ang, out_size, bilinear=(interpolation == BILINEAR),
                               bicubic=(interpolation == BICUBIC))
        out[k] = p.reshape(-1)
    return out


def contrast_avg_std(patches, obj_avg, obj_std):
    """cuicuilco ``contrast_enhance="AgeContrastEnhancement_Avg_Std"`` (``face_analysis.py:1042-1045, 1238``).
    The cuicuilco source is unavailable -- PARITY UNPINNED; defined here per patch as
    ``v = x / 255; y = (v - mean(v)) / (std(v) + 1e-8) * obj_std + obj_avg`` (float, unclipped)."""
    v = np.asarray(patches, dtype=np.float64) / 255.0
    m = v.mean(axis=1, keepdims=True)
    sd = v.std(axis=1, keepdims=True)
    return (v - m) / (sd + 1e-8) * obj_std + obj_avg
