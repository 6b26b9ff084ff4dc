"""Oracle (test infrastructure): the age-stage crop -- ``normalize_image`` + the 96 x 96 sub-sampling.

Follows ``face_normalization_tools.py:23-47, 111-329`` (called from ``face_analysis.py:1212-1222`` with
``normalization_method="eyes_inferred-mouth_areaZ"``, ``centering_mode="mid_eyes_inferred-mouth"``,
``rotation_mode="EyeLineRotation"``, ``integer_rotation_center=True``, ``out_size=(256, 260)``,
``allow_random_background=False``) statement by statement for the scalar geometry, and restates the three Pillow
resampling steps in numpy:

1. ``im.transform(crop_size, Image.EXTENT, crop_coordinates)`` -- default filter NEAREST, integer origin, unit scale:
   an exact copy of the window around the integer rotation centre, zero outside the image;
2. ``image_loader.rotate_improved(im_crop_first, rotation_angle, Image.BICUBIC)`` -- cuicuilco, UN-VENDORED.  Defined
   here as Pillow's own ``Image.rotate(angle, Image.BICUBIC)`` (rotation about ``(w / 2, h / 2)`` in continuous
   coordinates = the geometric centre of the pixel grid, same size, zero fill), which is what a helper that fixes
   the old PIL centre bug amounts to on a current Pillow.  PARITY UNPINNED for that one choice; the restatement of
   Pillow's rotate itself (matrix rounded to 15 digits, ``a0 (x + .5) + a1 (y + .5) + a2``, ``bicubic_filter8``) is PINNED
   against Pillow 12.2 live (``tests/test_oracle_normalize.py``);
3. ``im_rotated.transform(out_size, Image.EXTENT, (x0, y0, x1, y1), Image.BICUBIC)`` -- PINNED like ``oracle/crop.py``.

Then ``load_image_data_monoprocessor`` (cuicuilco, UN-VENDORED) with the arguments of ``face_analysis.py:1230-1246``:
96 x 96 samples at 1.9 image pixels per sample (``1.14 * 160 / 96``) starting at row ``260 / 2 - 96 * 1.9 / 2``, column
``256 / 2 - 96 * 1.9 / 2``, translated by ``(0, -6 / (160 / 96))`` *sampled* pixels (``trans_sampled=True``), contrast
``AgeContrastEnhancement_Avg_Std`` with ``obj_avg 0.0``, ``obj_std 0.16``.  Defined here as the EXTENT transform of that
box with the pipeline's interpolation (NEAREST) followed by ``oracle.crop.contrast_avg_std`` -- PARITY UNPINNED.
"""
import math

import numpy as np

from . import crop as ocrop

OUT_SIZE = (256, 260)
AGE_SAMPLING = 1.14 * (160.0 / 96)
AGE_OBJ_AVG, AGE_OBJ_STD = 0.0, 0.16


def approximate_mouth(eyes):
    """``compute_approximate_mouth_coordinates`` (``face_normalization_tools.py:23-47``)."""
    elx, ely, erx, ery = eyes
    eye_dx = erx - elx
    eye_dy = ery - ely
    mx = (erx + elx) / 2.0
    my = (ery + ely) / 2.0
    return mx - (42.0 / 37.0) * eye_dy, my + (42.0 / 37.0) * eye_dx


def geometry(eyes, im_width, im_height, out_size=OUT_SIZE):
    """Scalar part of ``normalize_image`` (``:150-311``): returns a dict with the integer crop window, the rotation
    angle and the EXTENT box of the last resampling step."""
    LeftEyeCenter_x, LeftEyeCenter_y, RightEyeCenter_x, RightEyeCenter_y = [float(v) for v in eyes]
    eyes_x_m = (RightEyeCenter_x + LeftEyeCenter_x) / 2.0
    eyes_y_m = (RightEyeCenter_y + LeftEyeCenter_y) / 2.0
    dist_eyes = np.sqrt((LeftEyeCenter_x - RightEyeCenter_x) ** 2 + (LeftEyeCenter_y - RightEyeCenter_y) ** 2)
    eye_line_angle = math.atan2(RightEyeCenter_y - LeftEyeCenter_y, RightEyeCenter_x - LeftEyeCenter_x) * 180 / math.pi
    if LeftEyeCenter_x > RightEyeCenter_x:
        raise ValueError("Warning: the eyes are ordered incorrectly!!!")           # the reference calls exit()
    desired_area = 37.0 * 42.0 / 2.0 * (37.5 / 37.0) ** 2
    inferred_mouth_x, inferred_mouth_y = approximate_mouth((LeftEyeCenter_x, LeftEyeCenter_y, RightEyeCenter_x, RightEyeCenter_y))
    height_triangle_with_inferred_mouth = np.sqrt((eyes_x_m - inferred_mouth_x) ** 2 + (eyes_y_m - inferred_mouth_y) ** 2)
    current_area_with_inferred_mouth = dist_eyes * height_triangle_with_inferred_mouth / 2.0
    midpoint_eyes_inferred_mouth_x = (eyes_x_m + inferred_mouth_x) / 2.0
    midpoint_eyes_inferred_mouth_y = (eyes_y_m + inferred_mouth_y) / 2.0
    # normalization_method == "eyes_inferred-mouth_areaZ"
    scale_factor = np.sqrt(current_area_with_inferred_mouth / desired_area)
    ori_width = out_size[0] * scale_factor / 2
    ori_height = out_size[1] * scale_factor / 2
    # centering_mode == "mid_eyes_inferred-mouth", rotation_mode == "EyeLineRotation", integer_rotation_center
    rotation_center_x = midpoint_eyes_inferred_mouth_x
    rotation_center_y = midpoint_eyes_inferred_mouth_y
    rotation_angle = eye_line_angle
    rotation_center_x_int = int(rotation_center_x + 0.5)
    rotation_center_y_int = int(rotation_center_y + 0.5)
    rotation_window_width = 2 * max(im_width - 1 - rotation_center_x_int + 0.5, rotation_center_x_int + 0.5)
    rotation_window_height = 2 * max(im_height - 1 - rotation_center_y_int + 0.5, rotation_center_y_int + 0.5)
    Delta_x = rotation_center_x - rotation_center_x_int
    Delta_y = rotation_center_y - rotation_center_y_int
    rotation_angle_rad = -rotation_angle * np.pi / 180.0
    Delta_x_rotated = Delta_x * np.cos(rotation_angle_rad) - Delta_y * np.sin(rotation_angle_rad)
    Delta_y_rotated = Delta_y * np.cos(rotation_angle_rad) + Delta_x * np.sin(rotation_angle_rad)
    rotation_crop_x0 = rotation_center_x_int - (rotation_window_width - 1) / 2.0
    rotation_crop_y0 = rotation_center_y_int - (rotation_window_height - 1) / 2.0
    rotation_crop_x1 = rotation_center_x_int + (rotation_window_width - 1) / 2.0 + 1
    rotation_crop_y1 = rotation_center_y_int + (rotation_window_height - 1) / 2.0 + 1
    crop_size = (int(rotation_window_width + 0.5), int(rotation_window_height + 0.5))
    center_x_geometric = (crop_size[0] - 1) / 2.0
    center_y_geometric = (crop_size[1] - 1) / 2.0
    new_center_x = center_x_geometric + Delta_x_rotated
    new_center_y = center_y_geometric + Delta_y_rotated
    x0 = (new_center_x - (ori_width - 1) / 2.0)
    x1 = (new_center_x + (ori_width - 1) / 2.0) + 1
    y0 = (new_center_y - (ori_height - 1) / 2.0)
    y1 = (new_center_y + (ori_height - 1) / 2.0) + 1
    return dict(crop_size=crop_size, crop_coordinates=(rotation_crop_x0, rotation_crop_y0, rotation_crop_x1, rotation_crop_y1),
                rotation_angle=rotation_angle, transform_coords=(x0, y0, x1, y1), scale_factor=scale_factor)


def rotate_matrix(angle_deg, w, h):
    """The affine matrix of Pillow's ``Image.rotate(angle)`` (``Image.py``: centre ``(w / 2, h / 2)``, no translation,
    coefficients rounded to 15 digits) or None for the angles Pillow short-cuts (0 -> copy)."""
    angle = angle_deg % 360.0
    if angle == 0:
        return None
    center = (w / 2, h / 2)
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-center[0]) + m[1] * (-center[1]) + m[2]
    m[5] = m[3] * (-center[0]) + m[4] * (-center[1]) + 0.0
    m[2] += center[0]
    m[5] += center[1]
    return m


def rotate_bicubic(img, angle_deg):
    """Pillow ``Image.rotate(angle, Image.BICUBIC)`` on a uint8 'L' image (same size, zero fill)."""
    H, W = img.shape
    m = rotate_matrix(angle_deg, W, H)
    if m is None:
        return img.copy()
    if angle_deg % 360.0 == 180:
        return img[::-1, ::-1].copy()                     # Pillow: transpose(ROTATE_180)
    xin = np.arange(W, dtype=np.float64) + 0.5
    yin = np.arange(H, dtype=np.float64) + 0.5
    X = (m[0] * xin[None, :] + m[1] * yin[:, None]) + m[2]
    Y = (m[3] * xin[None, :] + m[4] * yin[:, None]) + m[5]
    return ocrop._bicubic_sample(img, X, Y)


def crop_integer(img, crop_coordinates, crop_size):
    """``im.transform(crop_size, EXTENT, crop_coordinates)`` with the default NEAREST filter: integer origin, unit scale."""
    H, W = img.shape
    return ocrop.extent_nearest(img, crop_coordinates, crop_size)


def normalize_image(img, eyes, out_size=OUT_SIZE):
    """uint8 (260, 256): the reference's ``im2`` for the face whose eyes are (left x, left y, right x, right y)."""
    H, W = img.shape
    g = geometry(eyes, W, H, out_size)
    first = crop_integer(img, g["crop_coordinates"], g["crop_size"])
    rotated = rotate_bicubic(first, g["rotation_angle"])
    return ocrop.extent_bicubic(rotated, g["transform_coords"], out_size)


def age_box(age_subimage_width=96, age_subimage_height=96, out_size=OUT_SIZE):
    """Sampling box of the age patch inside the normalised image (``face_analysis.py:1183-1199``)."""
    age_image_width, age_image_height = out_size
    reduction_factor = 160.0 / 96
    age_sampling = 1.14 * reduction_factor
    first_row = age_image_height / 2.0 - age_subimage_height * age_sampling / 2.0
    first_column = age_image_width / 2.0 - age_subimage_width * age_sampling / 2.0
    tx = 0.0 / reduction_factor
    ty = -6.0 / reduction_factor
    x0 = first_column + tx * age_sampling           # trans_sampled=True: translations are in sampled pixels
    y0 = first_row + ty * age_sampling
    return (x0, y0, x0 + age_subimage_width * age_sampling, y0 + age_subimage_height * age_sampling)


def age_subimages(img, detections, age_subimage_width=96, age_subimage_height=96, interpolation=ocrop.NEAREST):
    """``age_subimages_arr`` of every detection row ([.., eye_l_x, eye_l_y, eye_r_x, eye_r_y, conf], columns 5..8):
    (N, 96 * 96) float64, contrast-normalised."""
    det = np.asarray(detections, dtype=np.float64).reshape(-1, 10)
    out = np.zeros((len(det), age_subimage_width * age_subimage_height))
    box = age_box(age_subimage_width, age_subimage_height)
    for j, row in enumerate(det):
        im2 = normalize_image(img, row[5:9])
        p = ocrop.extract_subimages(im2, np.asarray(box)[None, :], None, (age_subimage_width, age_subimage_height), interpolation)
        out[j] = ocrop.contrast_avg_std(p, AGE_OBJ_AVG, AGE_OBJ_STD)[0]
    return out
