"""Oracle (test infrastructure): the reference's per-image detection loop, stages 0 .. num_networks-6.

Follows ``FaceDetectUpdated.py:589-761`` statement by statement (per sampling value: grid, stage loop with
the skip rules of ``:674-682``, regression, coordinate update, discard mask, boolean compaction of coords /
angles / indices / sl / subimages_arr / confidence) using the oracle's own crop, flow, head and controller.
After the stage loop the reference refines the eyes with the eye networks (``:946-1041``); like the product
 this oracle stops at the approximate eye positions of
``compute_approximate_eye_boxes_coordinates`` (``face_analysis.py:61-135``) and then purges (``:1180``).
"""
import numpy as np

from . import controller as ctl
from . import crop as ocrop
from . import gauss as ogauss
from . import grid as ogrid
from . import nodes as onodes

CUT_OFFS_FACE = [0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.05]


def detect_image(image, header, network_types, networks, classifiers, smallest_face, num_face_stages,
                 cut_offs_face=CUT_OFFS_FACE, tol_posxy=1.1, tol_scale=1.1, tol_angle=1.1, interpolation=ocrop.NEAREST,
                 flow_execute=onodes.flow_execute, regression=ogauss.regression):
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
        for j, box in enumerate(curr_coords):
            eyes, _, _ = ctl.eye_boxes(box, rot_angle=curr_angles[j])
            detections.append(np.array([box[0], box[1], box[2], box[3], curr_angles[j], eyes[0], eyes[1], eyes[2],
                                        eyes[3], curr_confidence[j]]))
    raw = np.array(detections).reshape(-1, 10)
    purged = np.array(ctl.purge(detections)).reshape(-1, 10) if len(detections) else np.zeros((0, 10))
    return purged, dict(stage_counts=stage_counts, raw=raw)
