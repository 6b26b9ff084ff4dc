"""Oracle (test infrastructure): the reference's per-image detection loop, stages 0 .. num_networks-6.

Follows ``FaceDetectUpdated.py:589-761`` statement by statement (per sampling value: grid, stage loop with
the skip rules of ``:674-682``, regression, coordinate update, discard mask, boolean compaction of coords /
angles / indices / sl / subimages_arr / confidence) using the oracle's own crop, flow, head and controller.
After the stage loop the eyes are refined with the eye networks (``FaceDetectUpdated.py:946-1041`` ->
``find_Left_Right_eyes``, ``face_analysis.py:1036-1109``) when they are supplied, otherwise the approximate eye
positions of ``compute_approximate_eye_boxes_coordinates`` (``face_analysis.py:61-135``) are kept; then the purge
(``:1180``).  The age / race / gender stage (``face_analysis.py:1170-1306``) is not restated.

``curr_confidence`` is compacted at every stage here; the reference only re-assigns it at Disc stages
(``FaceDetectUpdated.py:758-759``), which is equivalent because the last face stage is a Disc.
"""
import numpy as np

from . import controller as ctl
from . import crop as ocrop
from . import gauss as ogauss
from . import grid as ogrid
from . import nodes as onodes

CUT_OFFS_FACE = [0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.05]


def find_eyes(image, curr_angles, eyes_box_orig, eye_header, eye_net, clf_x, clf_y, interpolation=ocrop.NEAREST,
              flow_execute=onodes.flow_execute, regression=ogauss.regression, tolerance_xy_eye=9.0):
    """``find_Left_Right_eyes`` with left_eye=1 (both eyes go through it unmirrored, ``face_analysis.py:1022-1033``)."""
    _, _, _, _, ew, eh, erw, erh = eye_header
    box = eyes_box_orig.copy()
    too_far = np.zeros(len(box), dtype=bool)
    if len(box) == 0:
        return box, too_far
    patches = ocrop.extract_subimages(image, box, curr_angles, (ew, eh), interpolation)
    patches = ocrop.contrast_avg_std(patches, 0.11, 0.15)
    sl = flow_execute(eye_net, patches)          # the reference executes the same flow once per label
    reg_x = regression(clf_x, sl[:, 0:clf_x.input_dim], clf_x.avg_labels)
    reg_y = regression(clf_y, sl[:, 0:clf_y.input_dim], clf_y.avg_labels)
    with np.errstate(invalid="ignore"):
        too_far |= np.abs(reg_x) >= tolerance_xy_eye
        too_far |= np.abs(reg_y) >= tolerance_xy_eye
    reg_out_x = (reg_x / 2.3719) * np.abs(box[:, 2] - box[:, 0]) / erw
    reg_out_y = (reg_y / 2.3719) * np.abs(box[:, 3] - box[:, 1]) / erh
    rot = -1 * 1 * curr_angles * np.pi / 180
    dx = reg_out_x * np.cos(rot) - reg_out_y * np.sin(rot)
    dy = reg_out_y * np.cos(rot) + reg_out_x * np.sin(rot)
    box[:, 0] = box[:, 0] - 1 * dx
    box[:, 2] = box[:, 2] - 1 * dx
    box[:, 1] = box[:, 1] - dy
    box[:, 3] = box[:, 3] - dy
    return box, too_far


def detect_image(image, header, network_types, networks, classifiers, smallest_face, num_face_stages,
                 cut_offs_face=CUT_OFFS_FACE, tol_posxy=1.1, tol_scale=1.1, tol_angle=1.1, interpolation=ocrop.NEAREST,
                 flow_execute=onodes.flow_execute, regression=ogauss.regression, eye_header=None):
    net_Dx, net_Dy, net_Dang, net_mins, net_maxs, sw, sh, rw, rh = header
    im_height, im_width = image.shape
    stage_counts = np.zeros(num_face_stages, dtype=np.int64)
    detections = []
    for sampling_value in ogrid.compute_sampling_values(im_width, im_height, sw, sh, smallest_face, net_mins, net_maxs):
        posX, posY, pw, ph, max_Dx_diff, max_Dy_diff = ogrid.compute_posX_posY_values(
            im_width, im_height, sw, sh, rw, rh, sampling_value, net_Dx, net_Dy)
        min_scale_radio = net_mins / 0.825
        max_scale_radio = net_maxs / 0.825
        n_orig, orig_coords = ogrid.compute_subimage_coordinates_from_posX_posY_values(posX, posY, pw, ph)
        orig_angles = np.zeros(n_orig)
        base_side = np.sqrt(pw ** 2 + ph ** 2)
        curr_coords = orig_coords.copy()
        curr_angles = orig_angles.copy()
        curr_orig_index = np.arange(n_orig)
        curr_confidence = np.zeros(n_orig)
        subimages_arr = None
        sl = None
        for k in range(num_face_stages):
            network_type = network_types[k][0:-1]
            serial = int(network_types[k][-1])
            cut_off_face = cut_offs_face[serial]
            skip_image_extraction = 0
            skip_feature_extraction = 0
            if k > 0 and network_types[k - 1][0:-1] == "Disc":
                skip_image_extraction = 1
            if networks[k] is None:
                skip_image_extraction = 1
                skip_feature_extraction = 1
            if skip_image_extraction == 0:
                subimages_arr = ocrop.extract_subimages(image, curr_coords, curr_angles, (sw, sh), interpolation)
            if subimages_arr is None or len(subimages_arr) == 0:
                continue
            if skip_feature_extraction == 0:
                sl = flow_execute(networks[k], subimages_arr)
            stage_counts[k] += sl.shape[0]
            D = classifiers[k].input_dim
            reg_out = regression(classifiers[k], sl[:, 0:D], classifiers[k].avg_labels) if sl.shape[0] > 0 else np.zeros(0)
            curr_coords, curr_angles = ctl.update_coordinates(network_type, curr_coords, curr_angles, reg_out, rw, rh, 0.825)
            wrong = ctl.patches_to_discard(network_type, curr_coords, curr_angles, reg_out, base_side, curr_orig_index,
                                           orig_coords, orig_angles, max_Dx_diff, max_Dy_diff, tol_posxy,
                                           max_scale_radio, min_scale_radio, tol_scale, net_Dang, tol_angle, cut_off_face)
            keep = wrong == 0
            curr_coords = curr_coords[keep].copy()
            curr_angles = curr_angles[keep].copy()
            curr_orig_index = curr_orig_index[keep].copy()
            sl = sl[keep].copy()
            subimages_arr = subimages_arr[keep, :].copy()
            if network_type == "Disc":
                curr_confidence = reg_out[keep].copy()
            else:
                curr_confidence = curr_confidence[keep].copy()
        n_face = len(curr_coords)
        eyes_orig = np.zeros((n_face, 4))
        eyesL_box_orig = np.zeros((n_face, 4))
        eyesR_box_orig = np.zeros((n_face, 4))
        for i, box in enumerate(curr_coords):
            eyes_orig[i], eyesL_box_orig[i], eyesR_box_orig[i] = ctl.eye_boxes(box, rot_angle=curr_angles[i])
        eye_net = networks[num_face_stages] if eye_header is not None and len(networks) > num_face_stages else None
        if eye_net is not None:
            cx, cy = classifiers[num_face_stages], classifiers[num_face_stages + 1]
            eyesL_box, farL = find_eyes(image, curr_angles, eyesL_box_orig, eye_header, eye_net, cx, cy, interpolation,
                                        flow_execute, regression)
            eyesR_box, farR = find_eyes(image, curr_angles, eyesR_box_orig, eye_header, eye_net, cx, cy, interpolation,
                                        flow_execute, regression)
            too_far = farL | farR
            eyesL = (eyesL_box[:, 0:2] + eyesL_box[:, 2:4]) / 2.0
            eyesR = (eyesR_box[:, 0:2] + eyesR_box[:, 2:4]) / 2.0
            eyesL, eyesR = eyesL[too_far == 0], eyesR[too_far == 0]
            curr_coords, curr_angles = curr_coords[too_far == 0, :], curr_angles[too_far == 0]
            # reference quirk kept: curr_confidence is NOT filtered by eye_xy_too_far (FaceDetectUpdated.py:1011-1017),
            # so face j of the survivors reports the confidence of face j of the pre-eye-stage list (:1036-1041)
        else:
            eyesL, eyesR = eyes_orig[:, 0:2], eyes_orig[:, 2:4]
        for j, box in enumerate(curr_coords):
            detections.append(np.array([box[0], box[1], box[2], box[3], curr_angles[j], eyesL[j][0], eyesL[j][1],
                                        eyesR[j][0], eyesR[j][1], curr_confidence[j]]))
    raw = np.array(detections).reshape(-1, 10)
    purged = np.array(ctl.purge(detections)).reshape(-1, 10) if len(detections) else np.zeros((0, 10))
    return purged, dict(stage_counts=stage_counts, raw=raw)
