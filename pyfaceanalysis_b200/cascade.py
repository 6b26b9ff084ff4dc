"""Batched detection cascade on the GPU.

Replaces the per-image / per-scale / per-stage loop of the reference (``FaceDetectUpdated.py:589-761``): for
every sampling value it builds the window grid, and for each of the first ``num_networks - 5`` stages
crops the windows, runs the stage's flow (or reuses the previous features when the network is ``None``),
evaluates the Gaussian regression head, moves / rotates / rescales the boxes, discards windows and compacts
every per-window array.

Here the windows of ALL scales of ALL images of a batch travel through the stages together, resident on
the device: grid coordinates are uploaded once, crop -> flow -> head -> controller -> compaction are
kernels of ``libhgsfa.so``.  Windows are independent until the per-image purge, so batching changes nothing but
the order of evaluation, and a window a stage discards may simply stay in the batch (its results are never
read): the arrays are compacted -- an 8-byte survivor count read by the host -- only after a Disc stage
that still holds many windows (in practice once, after Disc1), and at the end.  The stable compaction
keeps the reference's order (image, scale, window) inside the batch.

After the face stages the eyes of every surviving face are refined with the eye network and its two heads
(``FaceDetectUpdated.py:946-1041`` -> ``find_Left_Right_eyes``, ``face_analysis.py:1036-1109``): rotated 64x64 crops
of the two eye boxes, per-patch contrast normalisation ("AgeContrastEnhancement_Avg_Std", obj_avg 0.11, obj_std
0.15 -- definition in oracle/crop.py), one execution of the eye flow per eye (the reference runs it once per label
on the same input), EyeLX / EyeLY regressions, |reg| >= 9 discards the face.  Without eye networks the detections
carry the approximate eye positions of ``compute_approximate_eye_boxes_coordinates``.

Front end: ``prescale`` (NEAREST resize to <= 1000 px, ``FaceDetectUpdated.py:551-559``) runs on the device; back end:
``format_detections`` writes the reference's result lines (``FaceDetectUpdated.py:1258-1278``).
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _lib, grid
from .pipeline import CUT_OFFS_FACE, DEFAULTS

_TYPE_CODE = {"Disc": 0, "PosX": 1, "PosY": 2, "PAng": 3, "Scale": 4}


def approximate_eye_coordinates(boxes, angles):
    """Vectorised ``compute_approximate_eye_boxes_coordinates`` (reference ``face_analysis.py:61-135``),
    leftscreen_on_left=True; returns (N,4) = left eye x, y, right eye x, y (same float64 operation order)."""
    x0, y0, x1, y1 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    fc_x = (x0 + x1) / 2.0
    fc_y = (y0 + y1) / 2.0
    eye_dx = (37.0 / 2.0) * (np.abs(x1 - x0) / 64.0) / (2 * 0.825)
    eye_dy = (42.0 / 2.0) * (np.abs(y1 - y0) / 64.0) / (2 * 0.825)
    rad = angles * np.pi / 180
    er_dx = eye_dx * np.cos(rad) - eye_dy * np.sin(rad)
    er_dy = eye_dy * np.cos(rad) + eye_dx * np.sin(rad)
    el_dx = (-1 * eye_dx) * np.cos(rad) - eye_dy * np.sin(rad)
    el_dy = eye_dy * np.cos(rad) + (-1 * eye_dx) * np.sin(rad)
    return np.stack([fc_x + el_dx, fc_y - el_dy, fc_x + er_dx, fc_y - er_dy], axis=1)


def approximate_eye_boxes(boxes, angles):
    """Eye centres plus the left / right eye boxes of ``compute_approximate_eye_boxes_coordinates``:
    returns (eyes (N,4), left boxes (N,4), right boxes (N,4))."""
    eyes = approximate_eye_coordinates(boxes, angles)
    bw = (np.abs(boxes[:, 2] - boxes[:, 0]) / (64.0 * 2 * 0.825)) * (64 * 2.3719 / 2)
    bh = bw + 0.0

    def box(cx, cy):
        return np.stack([cx - bw / 2.0, cy - bh / 2.0, cx + bw / 2.0, cy + bh / 2.0], axis=1)
    return eyes, box(eyes[:, 0], eyes[:, 1]), box(eyes[:, 2], eyes[:, 3])


def group_confidences(cf, ok, grp):
    """The reference's confidence bookkeeping of the eye stage (``FaceDetectUpdated.py:1011-1017, 1036-1041``):
    ``curr_confidence`` of a (image, scale) group is NOT filtered by ``eye_xy_too_far``, so survivor j of a group reports
    the confidence of that group's face j.  cf: (n,), ok: (n,) bool, grp: (n, k) group keys, equal keys consecutive.
    One gather instead of a Python loop over the faces (hundreds per batch)."""
    n = len(cf)
    if n == 0:
        return np.empty(0)
    grp = np.asarray(grp).reshape(n, -1)
    starts = np.flatnonzero(np.concatenate([[True], np.any(grp[1:] != grp[:-1], axis=1)]))
    k_ok = np.add.reduceat(np.asarray(ok, dtype=np.int64), starts)
    first = np.cumsum(k_ok) - k_ok                         # output position of each group's first survivor
    return np.asarray(cf)[np.repeat(starts - first, k_ok) + np.arange(int(k_ok.sum()))]


def purge_detections(det, weight_confidences_by_area=True):
    """``purgue_detected_faces_angles_eyes_confidence`` (reference ``face_analysis.py:186-221``); rows are
    [x0, y0, x1, y1, angle, eye_l_x, eye_l_y, eye_r_x, eye_r_y, confidence].  Dozens of rows: stays on the host."""
    det = np.asarray(det, dtype=np.float64).reshape(-1, 10)
    if len(det) <= 1:
        return det.copy()
    conf = det[:, -1]
    if weight_confidences_by_area:
        areas = ((det[:, 7] - det[:, 5]) ** 2 + (det[:, 8] - det[:, 6]) ** 2) ** 0.5
        weighted = (1.0 - conf) * areas
        weighted = weighted / weighted.max()
    else:
        weighted = conf.copy()
    det = det[np.argsort(weighted)[::-1], :]

    def rel_err(a, b):
        dl = np.sqrt(((b[0:2] - a[0:2]) ** 2).sum())
        dr = np.sqrt(((b[2:4] - a[2:4]) ** 2).sum())
        de = np.sqrt(((b[0:2] - b[2:4]) ** 2).sum())
        return max(dl, dr) / de

    del rel_err        # below: its vectorised form, the same operations in the same order for every pair
    # e[i, j] = rel_err(row_i, row_j) for all pairs at once, then the reference's greedy scan over plain floats
    eye = det[:, 5:9]
    dl = np.sqrt(((eye[None, :, 0:2] - eye[:, None, 0:2]) ** 2).sum(axis=2))
    dr = np.sqrt(((eye[None, :, 2:4] - eye[:, None, 2:4]) ** 2).sum(axis=2))
    de = np.sqrt(((eye[:, 0:2] - eye[:, 2:4]) ** 2).sum(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        e = (np.maximum(dl, dr) / de[None, :]).tolist()
    unique = [0]
    for i in range(len(det)):
        row_e = e[i]
        min_d = 10000
        for j in unique:
            if row_e[j] < min_d:          # NaN never replaces the running minimum, as in the reference
                min_d = row_e[j]
        if min_d > 0.25:
            unique.append(i)
    return det[unique].copy()


def format_detections(det, right_screen_eye_first=False, attributes=None):
    """Result lines of one image exactly as ``FaceDetectUpdated.py:1258-1278`` writes them:
    ``"%d, %d, %d, %d, %f, %d, %d, %d, %d"`` of the rounded box, the angle and the rounded eye coordinates (left / right
    swapped with ``right_screen_eye_first``), optionally ``", %2.1f, %s, %s, %f"`` of age, race, gender and confidence
    (``attributes`` = (ages, races, genders), ``write_age_race_gender_confidence``), then ``" \\n"``."""
    lines = []
    for j, row in enumerate(np.asarray(det, dtype=np.float64).reshape(-1, 10)):
        r = np.round(row[0:9])
        eyes = (r[7], r[8], r[5], r[6]) if right_screen_eye_first else (r[5], r[6], r[7], r[8])
        line = "%d, %d, %d, %d, %f, %d, %d, %d, %d" % ((r[0], r[1], r[2], r[3], row[4]) + eyes)
        if attributes is not None:
            line += ", %2.1f, %s, %s, %f" % (attributes[0][j], attributes[1][j], attributes[2][j], row[9])
        lines.append(line + " \n")
    return "".join(lines)


class FaceDetector(object):
    """The face stages of a pipeline (``network_types[:num_networks - 5]``) as one device-resident cascade.

    networks[i] is a ``GpuFlow`` or ``None`` ("None0": reuse the previous features), classifiers[i] a
    ``GpuGaussianClassifier`` -- exactly what ``pipeline.load_networks_from_pipeline`` returns.
    """

    def __init__(self, header_net, network_types, networks, classifiers, num_face_stages=None,
                 cut_offs_face=None, interpolation=_lib.NEAREST, device=0, header_eye=None, attributes=None, **overrides):
        import torch
        self.torch = torch
        self.header = tuple(header_net)
        n_stages = num_face_stages if num_face_stages is not None else len(network_types) - 5
        self.types = list(network_types[:n_stages])
        self.networks = list(networks[:n_stages])
        self.classifiers = list(classifiers[:n_stages])
        self.cut_offs = list(cut_offs_face if cut_offs_face is not None else CUT_OFFS_FACE)
        self.cfg = dict(DEFAULTS)
        self.cfg.update(overrides)
        self.interpolation = int(interpolation)
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        if self.networks and self.networks[0] is None:
            raise ValueError("the first stage needs a network")
        self._labels = {}
        # age / race / gender stage (attributes.AttributeEstimator built from the last three network / classifier pairs)
        self.attributes = attributes
        # windows above which a Disc stage is followed by a compaction (one host round trip); below it discarded windows
        # simply ride along
        self.lazy_threshold = int(self.cfg.get("lazy_threshold", os.environ.get("HGSFA_LAZY_THRESHOLD", 4096)))
        # eye stage: network_types[n_stages] = EyeLX, [n_stages + 1] = EyeLY (same flow file, two heads)
        self.header_eye = tuple(header_eye) if header_eye is not None else None
        self.eye_net = None
        if self.header_eye is not None and len(networks) > n_stages and networks[n_stages] is not None:
            self.eye_net = networks[n_stages]
            self.eye_clf_x, self.eye_clf_y = classifiers[n_stages], classifiers[n_stages + 1]

    # ------------------------------------------------------------------------------------------
    def _labels_dev(self, clf):
        t = self._labels.get(id(clf))
        if t is None:
            t = self.torch.as_tensor(np.ascontiguousarray(clf.avg_labels, dtype=np.float64), device=self.dev)
            self._labels[id(clf)] = t
        return t

    def _regress(self, clf, sl, n, sp):
        """classifiers[i].regression(sl[:, 0:D], avg_labels) on device tensors -> float64 tensor (n,)."""
        torch = self.torch
        if sl.shape[1] < clf.input_dim:
            raise ValueError("x has dimension %d, should be %d" % (sl.shape[1], clf.input_dim))
        reg = torch.empty(n, dtype=torch.float64, device=self.dev)
        _lib.check(_lib.load().hgsfa_gauss_regress_device(
            clf.handle, C.c_void_p(sl.data_ptr()), _lib.F32, n, sl.stride(0), C.c_void_p(self._labels_dev(clf).data_ptr()),
            C.c_void_p(reg.data_ptr()), None, None, None, sp))
        return reg

    def _find_eyes(self, img_ptrs, img_hw, img_idx, angles_dev, angles_h, eye_boxes_h, sp):
        """``find_Left_Right_eyes`` (left_eye=1) for one eye of every surviving face."""
        torch = self.torch
        lib = _lib.load()
        _, _, _, _, ew, eh, erw, erh = self.header_eye
        n = len(eye_boxes_h)
        boxes = torch.as_tensor(np.ascontiguousarray(eye_boxes_h), device=self.dev)
        n_pad = (n + _lib.TILE - 1) // _lib.TILE * _lib.TILE
        patches = torch.empty(n_pad * ew * eh, dtype=torch.float32, device=self.dev)
        _lib.check(lib.hgsfa_crop_extent_batch_device(
            C.c_void_p(img_ptrs.data_ptr()), C.c_void_p(img_hw.data_ptr()), C.c_void_p(img_idx.data_ptr()),
            C.c_void_p(boxes.data_ptr()), C.c_void_p(angles_dev.data_ptr()), n, ew, eh, self.interpolation,
            C.c_void_p(patches.data_ptr()), _lib.F32, _lib.TILED, sp))
        _lib.check(lib.hgsfa_contrast_avg_std_device(C.c_void_p(patches.data_ptr()), n, ew * eh, 0.11, 0.15, sp))
        sl = self.eye_net.execute_torch(patches, layout=_lib.TILED, n=n)
        reg_x = self._regress(self.eye_clf_x, sl, n, sp).cpu().numpy()
        reg_y = self._regress(self.eye_clf_y, sl, n, sp).cpu().numpy()
        box = eye_boxes_h.copy()
        with np.errstate(invalid="ignore"):
            too_far = (np.abs(reg_x) >= 9.0) | (np.abs(reg_y) >= 9.0)
        reg_out_x = (reg_x / 2.3719) * np.abs(box[:, 2] - box[:, 0]) / erw
        reg_out_y = (reg_y / 2.3719) * np.abs(box[:, 3] - box[:, 1]) / erh
        rot = -1 * 1 * angles_h * np.pi / 180
        dx = reg_out_x * np.cos(rot) - reg_out_y * np.sin(rot)
        dy = reg_out_y * np.cos(rot) + reg_out_x * np.sin(rot)
        box[:, 0] = box[:, 0] - 1 * dx
        box[:, 2] = box[:, 2] - 1 * dx
        box[:, 1] = box[:, 1] - dy
        box[:, 3] = box[:, 3] - dy
        return box, too_far

    def _pyramid(self, im_width, im_height, smallest_face):
        """Window pyramid of one image size (host arrays + device copies), cached: it depends on the size only."""
        key = (int(im_width), int(im_height), float(smallest_face))
        cache = self.__dict__.setdefault("_pyramid_cache", {})
        p = cache.get(key)
        if p is None:
            p = grid.window_pyramid(im_width, im_height, self.header, smallest_face,
                                    self.cfg["patch_overlap_sampling"], self.cfg["patch_overlap_posx_posy"])
            p["coords_dev"] = self.torch.as_tensor(p["coords"], device=self.dev)
            p["patch_wh_dev"] = self.torch.as_tensor(p["patch_wh"], device=self.dev)
            if len(cache) > 64:
                cache.clear()
            cache[key] = p
        return p

    def prescale(self, images, prescale_size=None):
        """The reference's NEAREST prescale to at most ``prescale_size`` pixels per side (``FaceDetectUpdated.py:551-559``:
        ``factor = max(w / 1000., h / 1000.)``; if > 1: ``resize((int(w / factor), int(h / factor)), Image.NEAREST)``),
        on the device: Pillow's NEAREST resize is the EXTENT transform of the whole image, i.e. one window per image
        for the crop kernel.  images: 2-D uint8 numpy arrays or CUDA tensors; returns a list of CUDA uint8 tensors."""
        torch = self.torch
        lib = _lib.load()
        size = float(self.cfg["prescale_size"] if prescale_size is None else prescale_size)
        out = [None] * len(images)
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream(self.dev).cuda_stream
            sp = C.c_void_p(stream) if stream else None
            groups = {}                                        # images of one size are resized by ONE launch
            for k, im in enumerate(images):
                t = im if torch.is_tensor(im) else torch.as_tensor(np.ascontiguousarray(im, dtype=np.uint8), device=self.dev)
                h, w = int(t.shape[0]), int(t.shape[1])
                factor = max(w * 1.0 / size, h * 1.0 / size)
                if factor > 1.0:
                    groups.setdefault((h, w, int(w / factor), int(h / factor)), []).append((k, t))
                else:
                    out[k] = t
            for (h, w, pw, ph), items in groups.items():
                if pw > 1024 or ph > 1024:
                    raise ValueError("prescaled size %dx%d exceeds the crop kernel's 1024-pixel patch limit" % (pw, ph))
                n = len(items)
                key = ("prescale", h, w, n)
                cache = self.__dict__.setdefault("_prescale_cache", {})
                if key not in cache:
                    cache[key] = (torch.tensor([[0.0, 0.0, float(w), float(h)]] * n, dtype=torch.float64, device=self.dev),
                                  torch.tensor([[h, w]] * n, dtype=torch.int32, device=self.dev),
                                  torch.arange(n, dtype=torch.int32, device=self.dev))
                boxes, hw, idx = cache[key]
                ptrs = torch.tensor([t.data_ptr() for _, t in items], dtype=torch.int64, device=self.dev)
                dst = torch.empty((n, ph, pw), dtype=torch.uint8, device=self.dev)
                _lib.check(lib.hgsfa_crop_extent_batch_device(
                    C.c_void_p(ptrs.data_ptr()), C.c_void_p(hw.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(boxes.data_ptr()),
                    None, n, pw, ph, _lib.NEAREST, C.c_void_p(dst.data_ptr()), _lib.U8, _lib.ROWMAJOR, sp))
                for j, (k, _) in enumerate(items):
                    out[k] = dst[j]
        return out

    def detect(self, images, smallest_face=0.2, return_trace=False, benchmark=None, estimate_attributes=False):
        """images: list of 2-D uint8 arrays (the reference's mode-'L' image, already prescaled) or CUDA uint8 tensors
        (e.g. from ``prescale``).  Returns a list (one entry per image) of (M,10) float64 detection arrays after the
        purge; with ``return_trace`` also a dict with the per-stage window counts and the un-purged detections.
        ``benchmark``: an object with the reference's ``Benchmark.add_task_ellapsed(label, seconds)``
        (``benchmarking.py:39``); it receives the device time of every phase under the reference's labels
        (``FaceDetectUpdated.py:691,711,724,760``), summed over the stages of the batch.
        ``estimate_attributes`` (needs ``attributes=`` at construction): also run the age / race / gender stage on the
        purged detections (``FaceDetectUpdated.py:1187`` -> ``estimate_age_race_gender``); the per-image dicts come back
        as a second return value, or as ``trace["attributes"]`` with ``return_trace``."""
        # the detector's device becomes current for the call: torch allocations, the stream looked up below and
        # every kernel launch then agree, whatever device the calling thread had selected
        with self.torch.cuda.device(self.dev):
            return self._detect(images, smallest_face, return_trace, benchmark, estimate_attributes)

    def _detect(self, images, smallest_face, return_trace, benchmark, estimate_attributes=False):
        torch = self.torch
        lib = _lib.load()
        dev = self.dev
        net_Dx, net_Dy, net_Dang, net_mins, net_maxs, sw, sh, rw, rh = self.header
        stream = torch.cuda.current_stream(dev).cuda_stream
        sp = C.c_void_p(stream) if stream else None
        prof = {} if os.environ.get("HGSFA_DETECT_PROFILE") else None      # phase -> seconds (synchronising: diagnostics only)
        t_last = [time.perf_counter()]
        marks = [] if benchmark is not None else None                      # (label, event): device times, read once at the end

        def mark(name, label=None):
            if marks is not None and label is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((label, ev))
            if prof is not None:
                torch.cuda.synchronize(dev)
                now = time.perf_counter()
                prof[name] = prof.get(name, 0.0) + now - t_last[0]
                t_last[0] = now

        mark("start", "")
        # ---- window pyramid of every image, one batch ----
        shapes = [(int(im.shape[0]), int(im.shape[1])) for im in images]
        pyr = [self._pyramid(w, h, smallest_face) for h, w in shapes]
        n0 = int(sum(len(p["coords"]) for p in pyr))
        n_stages = len(self.types)
        if n0 == 0:
            out = [np.zeros((0, 10)) for _ in images]
            tr = dict(stage_counts=np.zeros(n_stages, dtype=np.int64), raw=[np.zeros((0, 10)) for _ in images], n_windows=0,
                      disc_scores={}, host_syncs=0)
            return (out, tr) if return_trace else out
        scale_h = np.concatenate([p["scale"] for p in pyr])

        imgs_dev = [im if torch.is_tensor(im) else torch.as_tensor(np.ascontiguousarray(im, dtype=np.uint8), device=dev)
                    for im in images]
        for t in imgs_dev:
            if t.dtype != torch.uint8 or t.dim() != 2 or not t.is_contiguous() or t.device != dev:
                raise ValueError("images must be contiguous 2-D uint8 arrays (mode 'L') on %s" % (dev,))
        img_ptrs = torch.tensor([t.data_ptr() for t in imgs_dev], dtype=torch.int64, device=dev)
        img_hw = torch.tensor([[t.shape[0], t.shape[1]] for t in imgs_dev], dtype=torch.int32, device=dev)

        # per-window arrays are replicated on the device from the cached per-size pyramids: no per-window upload
        orig_coords = torch.cat([p["coords_dev"] for p in pyr]) if len(pyr) > 1 else pyr[0]["coords_dev"].clone()
        orig_angles = torch.zeros(n0, dtype=torch.float64, device=dev)
        patch_wh = torch.cat([p["patch_wh_dev"] for p in pyr]) if len(pyr) > 1 else pyr[0]["patch_wh_dev"]
        coords = orig_coords.clone()
        angles = orig_angles.clone()
        img_idx = torch.repeat_interleave(torch.arange(len(pyr), dtype=torch.int32, device=dev),
                                          torch.tensor([len(p["coords"]) for p in pyr], device=dev), output_size=n0)
        orig_idx = torch.arange(n0, dtype=torch.int32, device=dev)
        conf = torch.zeros(n0, dtype=torch.float64, device=dev)
        keep = torch.empty(n0, dtype=torch.uint8, device=dev)
        alive = torch.ones(n0, dtype=torch.uint8, device=dev)              # windows not discarded by any stage so far
        src_index = torch.empty(n0, dtype=torch.int32, device=dev)
        count_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        counts_dev = torch.zeros(n_stages, dtype=torch.int64, device=dev)  # windows entering every stage, read once
        scratch = torch.empty(max(1, (n0 + 1023) // 1024), dtype=torch.int32, device=dev)
        sl = None
        patches = None                                                    # (n, sw * sh) uint8, row-major
        n = n0
        disc_scores = {}
        host_syncs = 0
        min_r = net_mins / 0.825
        max_r = net_maxs / 0.825

        def compact():
            """Order-preserving compaction of every per-window array by `alive`: the ONE kind of host round trip of the
            stage loop (the survivor count sizes the next launches)."""
            nonlocal coords, angles, img_idx, orig_idx, conf, sl, patches, alive, n, host_syncs
            _lib.check(lib.hgsfa_compact_index_device(C.c_void_p(alive.data_ptr()), n, C.c_void_p(src_index.data_ptr()),
                                                      C.c_void_p(count_dev.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                                      scratch.numel(), sp))
            n_new = int(count_dev.item())
            host_syncs += 1
            if n_new == n:
                return

            def gather(t, row_bytes):
                out = torch.empty((n_new,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                if n_new:
                    _lib.check(lib.hgsfa_gather_rows_device(C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()),
                                                            C.c_void_p(src_index.data_ptr()), n_new, row_bytes, sp))
                return out
            coords = gather(coords, 32)
            angles = gather(angles, 8)
            img_idx = gather(img_idx, 4)
            orig_idx = gather(orig_idx, 4)
            conf = gather(conf, 8)
            if sl is not None:
                sl = gather(sl.contiguous(), sl.shape[1] * 4)
            if patches is not None:
                patches = gather(patches, patches.shape[1])
            alive = torch.ones(n_new, dtype=torch.uint8, device=dev)
            n = n_new

        mark("pyramid+upload", "Window creation, and pre-computations")
        for k, full_type in enumerate(self.types):
            if n == 0:
                break
            counts_dev[k] = alive[:n].sum()
            ntype, serial = full_type[:-1], int(full_type[-1])
            net, clf = self.networks[k], self.classifiers[k]
            if net is not None:
                # the reference re-extracts unless the previous stage was a Disc, whose compacted subimages_arr it keeps
                # (FaceDetectUpdated.py:674-682); a Disc stage does not move boxes, so the kept patches are the same pixels
                reuse = k > 0 and self.types[k - 1][:-1] == "Disc" and patches is not None and patches.shape[0] == n
                if not reuse:
                    patches = torch.empty((n, sw * sh), dtype=torch.uint8, device=dev)
                    _lib.check(lib.hgsfa_crop_extent_batch_device(
                        C.c_void_p(img_ptrs.data_ptr()), C.c_void_p(img_hw.data_ptr()), C.c_void_p(img_idx.data_ptr()),
                        C.c_void_p(coords.data_ptr()), C.c_void_p(angles.data_ptr()), n, sw, sh, self.interpolation,
                        C.c_void_p(patches.data_ptr()), _lib.U8, _lib.ROWMAJOR, sp))
                mark("crop[%d]" % min(k, 1), "Extraction of subimages patches")
                sl = net.execute_torch(patches)
                mark("flow[%d]" % min(k, 1), "Feature extraction")
            elif sl is None:
                raise ValueError("stage %s reuses features but none were computed" % full_type)
            reg = self._regress(clf, sl, n, sp)
            mark("head[%d]" % min(k, 1), "Regression")
            if return_trace and ntype == "Disc":
                disc_scores[full_type] = (reg, alive.clone())
            params = np.array([net_Dx, net_Dy, net_Dang, rw, rh, min_r, max_r, self.cfg["tolerance_posxy_deviation"],
                               self.cfg["tolerance_scale_deviation"], self.cfg["tolerance_angle_deviation"], 0.825,
                               self.cut_offs[serial]], dtype=np.float64)
            _lib.check(lib.hgsfa_cascade_update_device(
                _TYPE_CODE[ntype], C.c_void_p(coords.data_ptr()), C.c_void_p(angles.data_ptr()),
                C.c_void_p(reg.data_ptr()), C.c_void_p(orig_coords.data_ptr()), C.c_void_p(orig_angles.data_ptr()),
                C.c_void_p(orig_idx.data_ptr()), C.c_void_p(patch_wh.data_ptr()), n, _lib.ptr(params),
                C.c_void_p(keep.data_ptr()), C.c_void_p(conf.data_ptr()) if ntype == "Disc" else None, sp))
            alive[:n] &= keep[:n]
            # Windows are independent, so a discarded window may stay in the batch (its results are never read): the arrays
            # are compacted -- the only host round trip -- where it pays, after a Disc stage while the batch is still
            # large (the first Disc stage removes most windows); later stages run on the small remainder unsynchronised.
            if ntype == "Disc" and n > self.lazy_threshold:
                compact()
            mark("update+compact[%d]" % min(k, 1), "Adjusted according to regression")
        compact()                                                          # survivors of the face stages
        counts = counts_dev.cpu().numpy()
        if return_trace:
            disc_scores = {name: r[a.bool()].cpu().numpy() for name, (r, a) in disc_scores.items()}

        # ---- survivors -> eyes -> detections (host arithmetic on dozens of rows, device compute for the eye flow) ----
        boxes = coords[:n].cpu().numpy()
        ang = angles[:n].cpu().numpy()
        im_of = img_idx[:n].cpu().numpy()
        cf = conf[:n].cpu().numpy()
        scale_of = scale_h[orig_idx[:n].cpu().numpy()] if n else np.zeros(0, dtype=np.int32)
        if n and self.eye_net is not None:
            _, boxL, boxR = approximate_eye_boxes(boxes, ang)
            # left and right eye patches of all faces in one batch (per-window work: identical results, half the launches)
            both_box, both_far = self._find_eyes(img_ptrs, img_hw, torch.cat([img_idx[:n], img_idx[:n]]),
                                                 torch.cat([angles[:n], angles[:n]]), np.concatenate([ang, ang]),
                                                 np.concatenate([boxL, boxR]), sp)
            eyesL_box, farL, eyesR_box, farR = both_box[:n], both_far[:n], both_box[n:], both_far[n:]
            ok = ~(farL | farR)
            eyes = np.concatenate([(eyesL_box[:, 0:2] + eyesL_box[:, 2:4]) / 2.0,
                                   (eyesR_box[:, 0:2] + eyesR_box[:, 2:4]) / 2.0], axis=1)[ok]
            # reference quirk kept (FaceDetectUpdated.py:1011-1017, 1036-1041): curr_confidence is not filtered by
            # eye_xy_too_far, so survivor j of a (scale, image) group reports the confidence of that group's face j
            cf_out = group_confidences(cf, ok, np.stack([im_of, scale_of], axis=1))
            boxes, ang, im_of, cf = boxes[ok], ang[ok], im_of[ok], cf_out
            n = len(boxes)
        else:
            eyes = approximate_eye_coordinates(boxes, ang) if n else np.zeros((0, 4))
        mark("eyes", "Eye localization (patches, feature extraction, regression)")
        raw = np.concatenate([boxes, ang[:, None], eyes, cf[:, None]], axis=1) if n else np.zeros((0, 10))
        per_image_raw = [raw[im_of == k] for k in range(len(images))]
        result = [purge_detections(r) if len(r) else np.zeros((0, 10)) for r in per_image_raw]
        mark("purge", "Purgued repeated face detections")
        attrs = None
        if estimate_attributes:
            if self.attributes is None:
                raise ValueError("estimate_attributes needs FaceDetector(..., attributes=AttributeEstimator(...))")
            attrs = self.attributes.estimate_detections(imgs_dev, result)
            mark("attributes", "Age/race/gender: normalized image, image array, feature extraction, regressions")
        if prof is not None:
            self.last_profile = prof
        if marks is not None:
            # device time between consecutive marks, booked under the reference's Benchmark labels
            torch.cuda.synchronize(dev)
            for (_, e_prev), (label, e) in zip(marks[:-1], marks[1:]):
                benchmark.add_task_ellapsed(label, e_prev.elapsed_time(e) * 1e-3)
        if return_trace:
            return result, dict(stage_counts=counts, raw=per_image_raw, n_windows=n0, disc_scores=disc_scores,
                                host_syncs=host_syncs, attributes=attrs)
        return (result, attrs) if estimate_attributes else result
