"""Window extraction facade: drop-in for ``load_network_subimages`` (reference ``face_analysis.py:775-800``).

The reference builds a list of PIL patches with cuicuilco's ``extract_subimages_rotate`` (one
``Image.transform(size, EXTENT, box, filter)`` per window) and stacks them with ``images_asarray`` into a
``float64 (N, w*h)`` array of 0..255 values.  Here one kernel pair (``csrc/crop.cu``) produces the whole
batch.
"""
from __future__ import annotations

import numpy as np

from . import _lib

NEAREST, BILINEAR, BICUBIC = _lib.NEAREST, _lib.BILINEAR, _lib.BICUBIC


def _image_array(image):
    if hasattr(image, "mode") and hasattr(image, "size"):   # PIL image, mode 'L' like the reference loads
        if image.mode != "L":
            image = image.convert("L")
        return np.ascontiguousarray(np.asarray(image, dtype=np.uint8))
    a = np.asarray(image)
    if a.ndim != 2 or a.dtype != np.uint8:
        raise TypeError("image must be a PIL 'L' image or a 2-D uint8 array")
    return np.ascontiguousarray(a)


def extract_subimages(image, coords, angles=None, out_size=(64, 64), interpolation=NEAREST,
                      out_dtype=np.float64, device=0):
    """``(N, ow*oh)`` patches, row-major, values 0..255 (``images_asarray`` layout)."""
    img = _image_array(image)
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 4)
    n = coords.shape[0]
    if angles is not None:
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        if angles.shape[0] != n:
            raise ValueError("%d angles for %d windows" % (angles.shape[0], n))
    ow, oh = int(out_size[0]), int(out_size[1])
    out = np.empty((n, ow * oh), dtype=out_dtype)
    if n == 0:
        return out
    _lib.check(_lib.load().hgsfa_crop_extent(_lib.ptr(img), img.shape[0], img.shape[1], _lib.ptr(coords),
                                             _lib.ptr(angles), n, ow, oh, int(interpolation), _lib.ptr(out),
                                             _lib.dtype_code(out.dtype), int(device), None))
    return out


def load_network_subimages(images, curr_image_indices, curr_subimage_coordinates, curr_angles, subimage_width,
                           subimage_height, interpolation_format=NEAREST, contrast_normalize=False, device=0):
    """Same signature and result as the reference function (``face_analysis.py:775``).  Like every call
    site of the reference (``FaceDetectUpdated.py:686-687``) ``contrast_normalize`` must be False."""
    if contrast_normalize:
        raise NotImplementedError("contrast_normalize=True is never used by the reference pipeline")
    idx = np.asarray(curr_image_indices).reshape(-1)
    coords = np.asarray(curr_subimage_coordinates, dtype=np.float64).reshape(-1, 4)
    if len(coords) == 0:
        return np.zeros((0, 0))
    out = np.empty((len(coords), subimage_width * subimage_height), dtype=np.float64)
    for im in np.unique(idx):
        sel = np.nonzero(idx == im)[0]
        out[sel] = extract_subimages(images[int(im)], coords[sel], None if curr_angles is None else
                                     np.asarray(curr_angles)[sel], (subimage_width, subimage_height),
                                     interpolation_format, np.float64, device)
    return out
