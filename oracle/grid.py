"""Oracle (test infrastructure): window-grid enumeration, restated from the reference source.

``compute_sampling_values``  -- ``face_analysis.py:575-607`` (adaptive_grid_scale branch, the default
                                ``FaceDetectUpdated.py:116``)
``compute_posX_posY_values`` -- ``face_analysis.py:610-657`` (adaptive_grid_coords branch)
``compute_subimage_coordinates_from_posX_posY_values`` -- ``face_analysis.py:661-669``

Python-2 semantics kept: ``numpy.ceil(...)`` was passed to ``linspace`` as ``num`` (``:640-646``), which
numpy of that era truncated to int.  All arithmetic is float64 in the reference's operation order, so
the result is bit-exact by construction (it IS the reference arithmetic, minus the Python loop).
"""
import numpy as np


def compute_sampling_values(im_width, im_height, subimage_width, subimage_height, smallest_face,
                            net_mins, net_maxs, patch_overlap_sampling=1.1, adaptive_grid_scale=True):
    min_side = min(im_height, im_width)
    min_box_side = max(20, min_side * smallest_face * 0.825 / net_mins)
    min_sampling_value = min_box_side * 1.0 / subimage_width
    if not adaptive_grid_scale:
        return [min_sampling_value]
    sampling_values = []
    sampling_value = min_sampling_value
    new_grid_step = (net_maxs / net_mins) / patch_overlap_sampling
    while (subimage_width * sampling_value * net_mins / 0.825 < im_width) and (
            subimage_height * sampling_value * net_mins / 0.825 < im_height):
        sampling_values.append(sampling_value)
        sampling_value *= new_grid_step
    return sampling_values


def compute_posX_posY_values(im_width, im_height, subimage_width, subimage_height, regression_width,
                             regression_height, sampling_value, net_Dx, net_Dy,
                             patch_overlap_posx_posy=1.1):
    patch_width = subimage_width * sampling_value
    patch_height = subimage_height * sampling_value
    patch_horizontal_separation = net_Dx * 2.0 * patch_width / regression_width
    patch_vertical_separation = net_Dy * 2.0 * patch_height / regression_height
    num_x_patches = np.ceil((1 + (im_width - patch_width) / patch_horizontal_separation) * patch_overlap_posx_posy)
    posX_values = np.linspace(0.0, im_width - patch_width, int(num_x_patches))
    num_y_patches = np.ceil((1 + (im_height - patch_height) / patch_vertical_separation) * patch_overlap_posx_posy)
    posY_values = np.linspace(0.0, im_height - patch_height, int(num_y_patches))
    max_Dx_diff = net_Dx * patch_width / regression_width
    max_Dy_diff = net_Dy * patch_height / regression_height
    return posX_values, posY_values, patch_width, patch_height, max_Dx_diff, max_Dy_diff


def compute_subimage_coordinates_from_posX_posY_values(posX_values, posY_values, patch_width, patch_height):
    orig_num_subimages = len(posX_values) * len(posY_values)
    coords = np.zeros((orig_num_subimages, 4))
    for j, posY in enumerate(posY_values):
        for i, posX in enumerate(posX_values):
            coords[j * len(posX_values) + i] = np.array(
                [posX, posY, posX + patch_width - 1, posY + patch_height - 1])
    return orig_num_subimages, coords


def prescaled_size(width, height, max_side=1000):
    """``FaceDetectUpdated.py:551-559``: shrink so that max(side) <= 1000 (default image_prescaling)."""
    prescaling_factor = max(width, height) * 1.0 / max_side
    if prescaling_factor > 1.0:
        return int(width / prescaling_factor), int(height / prescaling_factor), prescaling_factor
    return width, height, 1.0


def enumerate_windows(im_width, im_height, header, smallest_face):
    """All scales of one image: list of (sampling_value, coords(N,4), geometry dict)."""
    (net_Dx, net_Dy, net_Dang, net_mins, net_maxs, sw, sh, rw, rh) = header
    out = []
    for s in compute_sampling_values(im_width, im_height, sw, sh, smallest_face, net_mins, net_maxs):
        px, py, pw, ph, mdx, mdy = compute_posX_posY_values(im_width, im_height, sw, sh, rw, rh, s, net_Dx, net_Dy)
        n, coords = compute_subimage_coordinates_from_posX_posY_values(px, py, pw, ph)
        out.append((s, coords, dict(patch_width=pw, patch_height=ph, max_Dx_diff=mdx, max_Dy_diff=mdy,
                                    n_x=len(px), n_y=len(py))))
    return out
