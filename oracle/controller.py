"""Oracle (test infrastructure): cascade controller arithmetic, restated from the reference source.

``update_current_subimage_coordinates`` -- ``face_analysis.py:803-840``
``identify_patches_to_discard``          -- ``face_analysis.py:842-887``
``compute_approximate_eye_boxes_coordinates`` -- ``face_analysis.py:61-135``
``purgue_detected_faces_angles_eyes_confidence`` -- ``face_analysis.py:186-221``

The source of these functions is in the reference tree, so the float64 operation order below is the
reference's own; comparisons keep their strictness (``>``, ``>=``, ``<``) as written.
"""
import numpy as np


def update_coordinates(network_type, coords, angles, reg_out, regression_width, regression_height,
                       desired_sampling=0.825):
    """In place on ``coords`` like the reference; returns (coords, angles)."""
    if network_type == "Disc":
        pass
    elif network_type == "PosX":
        width = coords[:, 2] - coords[:, 0]
        reg = reg_out * width / regression_width
        coords[:, 0] = coords[:, 0] - reg
        coords[:, 2] = coords[:, 2] - reg
    elif network_type == "PosY":
        height = coords[:, 3] - coords[:, 1]
        reg = reg_out * height / regression_height
        coords[:, 1] = coords[:, 1] - reg
        coords[:, 3] = coords[:, 3] - reg
    elif network_type == "PAng":
        angles = angles + reg_out
    elif network_type == "Scale":
        old_width = coords[:, 2] - coords[:, 0]
        old_height = coords[:, 3] - coords[:, 1]
        x_center = (coords[:, 2] + coords[:, 0]) / 2.0
        y_center = (coords[:, 3] + coords[:, 1]) / 2.0
        width = old_width / reg_out * desired_sampling
        height = old_height / reg_out * desired_sampling
        coords[:, 0] = x_center - width / 2.0
        coords[:, 2] = x_center + width / 2.0
        coords[:, 1] = y_center - height / 2.0
        coords[:, 3] = y_center + height / 2.0
    else:
        raise Exception(("Network type unknown!!!: ", network_type))
    return coords, angles


def patches_to_discard(network_type, coords, angles, curr_disc, base_side, curr_orig_index, orig_coords,
                       orig_angles, max_Dx_diff, max_Dy_diff, tolerance_posxy_deviation, max_scale_radio,
                       min_scale_radio, tolerance_scale_deviation, net_Dang, tolerance_angle_deviation,
                       cut_off_face):
    if network_type == "PosX":
        deltas = (coords[:, 2] + coords[:, 0]) / 2 - \
                 (orig_coords[curr_orig_index][:, 2] + orig_coords[curr_orig_index][:, 0]) / 2
        return np.abs(deltas) > (max_Dx_diff * tolerance_posxy_deviation)
    if network_type == "PosY":
        deltas = (coords[:, 3] + coords[:, 1]) / 2 - \
                 (orig_coords[curr_orig_index][:, 3] + orig_coords[curr_orig_index][:, 1]) / 2
        return np.abs(deltas) > (max_Dy_diff * tolerance_posxy_deviation)
    if network_type == "PAng":
        return (angles > orig_angles[curr_orig_index] + net_Dang * tolerance_angle_deviation) | \
               (angles < orig_angles[curr_orig_index] - net_Dang * tolerance_angle_deviation)
    if network_type == "Scale":
        magnitudes = ((coords[:, 0:2] - coords[:, 2:4]) ** 2).sum(axis=1)
        sides = np.sqrt(magnitudes)
        return (sides / base_side > max_scale_radio * tolerance_scale_deviation) | \
               (sides / base_side < min_scale_radio / tolerance_scale_deviation)
    if network_type == "Disc":
        with np.errstate(invalid="ignore"):
            return curr_disc >= cut_off_face          # NaN >= c is False: the window is kept
    raise Exception("Unknown network type:" + str(network_type))


def eye_boxes(box, rot_angle=0.0):
    """``compute_approximate_eye_boxes_coordinates`` for one box (leftscreen_on_left=True)."""
    x0, y0, x1, y1 = box
    fc_x = (x0 + x1) / 2.0
    fc_y = (y0 + y1) / 2.0
    eye_dx = (37.0 / 2.0) * (np.abs(x1 - x0) / 64.0) / (2 * 0.825)
    eye_dy = (42.0 / 2.0) * (np.abs(y1 - y0) / 64.0) / (2 * 0.825)
    box_width = (np.abs(x1 - x0) / (64.0 * 2 * 0.825)) * (64 * 2.3719 / 2)
    box_height = box_width + 0.0
    rad = rot_angle * np.pi / 180
    er_dx = eye_dx * np.cos(rad) - eye_dy * np.sin(rad)
    er_dy = eye_dy * np.cos(rad) + eye_dx * np.sin(rad)
    el_dx = (-1 * eye_dx) * np.cos(rad) - eye_dy * np.sin(rad)
    el_dy = eye_dy * np.cos(rad) + (-1 * eye_dx) * np.sin(rad)
    el_x = fc_x + el_dx
    er_x = fc_x + er_dx
    el_y = fc_y - el_dy
    er_y = fc_y - er_dy
    left = np.array([el_x - box_width / 2.0, el_y - box_height / 2.0, el_x + box_width / 2.0, el_y + box_height / 2.0])
    right = np.array([er_x - box_width / 2.0, er_y - box_height / 2.0, er_x + box_width / 2.0, er_y + box_height / 2.0])
    return np.array([el_x, el_y, er_x, er_y]), left, right


def relative_error_detection(app_eye_coords, eye_coords):
    dist_left = np.sqrt(((eye_coords[0:2] - app_eye_coords[0:2]) ** 2).sum())
    dist_right = np.sqrt(((eye_coords[2:4] - app_eye_coords[2:4]) ** 2).sum())
    dist_eyes = np.sqrt(((eye_coords[0:2] - eye_coords[2:4]) ** 2).sum())
    return max(dist_left, dist_right) / dist_eyes


def purge(detections, weight_confidences_by_area=True):
    """``purgue_detected_faces_angles_eyes_confidence``: rows = [x0,y0,x1,y1,angle,elx,ely,erx,ery,conf]."""
    det = np.array(detections)
    if len(det) > 1:
        conf = det[:, -1]
        if weight_confidences_by_area:
            areas = ((det[:, 7] - det[:, 5]) ** 2 + (det[:, 8] - det[:, 6]) ** 2) ** 0.5
            weighted = (1.0 - conf) * areas
            weighted = weighted / weighted.max()
        else:
            weighted = conf.copy()
        ordering = np.argsort(weighted)[::-1]
        det = det[ordering, :]
        unique = [det[0]]
        for row in det:
            min_d = 10000
            for row2 in unique:
                err = relative_error_detection(row[5:9], row2[5:9])
                if err < min_d:
                    min_d = err
            if min_d > 0.25:
                unique.append(row)
        return unique
    return det.copy()
