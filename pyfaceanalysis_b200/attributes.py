"""Age / race / gender estimation on normalised face crops: the compute part of the reference's
``estimate_age_race_gender`` (``face_analysis.py:1170-1306``).

Reference, per final face::

    sl = networks[num_networks - 3].execute(age_subimages_arr)                       # (1, 9216) -> features
    age, age_std = classifiers[num_networks - 3].regression(sl[:, :D], avg_labels, estimate_std=True)
    race   = classifiers[num_networks - 2].regression(sl[:, :D], avg_labels)
    gender = classifiers[num_networks - 1].regression(sl[:, :D], avg_labels)
    gender_confidences = |gender|;  race_confidences = |race| / 2
    strings through map_real_gender_labels_to_strings / map_real_race_labels_to_strings

Here: one batched flow execute over all faces and three batched heads.  The 96 x 96 crop itself --
``normalize_image`` (``face_normalization_tools.py:111-329``: integer EXTENT crop, ``rotate_improved(BICUBIC)``, BICUBIC
EXTENT resample to 256 x 260) followed by ``load_image_data_monoprocessor`` (sub-sampling at 1.9 px / sample, contrast
enhancement) -- is one kernel over all faces (``csrc/crop.cu: age_crop_kernel``, geometry in ``normalize.py``) plus the
contrast kernel: ``estimate_detections`` goes from images + detection rows to the estimates; ``estimate`` still takes
the (N, 9216) patch matrix the reference calls ``age_subimages_arr``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, normalize


def map_real_gender_labels_to_strings(gender_label_array, long_text=True):
    """``face_analysis.py:333-351``: label <= 0 -> Male, else Female; |label| > 1.000001 is an error."""
    out = []
    for label in gender_label_array:
        out.append(("Male" if long_text else "M") if label <= 0 else ("Female" if long_text else "F"))
        if label < -1.000001 or label > 1.000001:
            raise Exception("Unrecognized label: " + str(label))
    return out


def map_real_race_labels_to_strings(race_label_array, long_text=True):
    """``face_analysis.py:354-371``: label <= 0 -> Black, else White; |label| > 2.000001 is an error."""
    out = []
    for label in race_label_array:
        out.append(("Black" if long_text else "B") if label <= 0.0 else ("White" if long_text else "W"))
        if label < -2.000001 or label > 2.000001:
            raise Exception("Unrecognized label: " + str(label))
    return out


class AttributeEstimator(object):
    """networks[-3] (a ``GpuFlow``) and classifiers[-3:] (``GpuGaussianClassifier``: age, race, gender) of a pipeline."""

    def __init__(self, network, clf_age, clf_race, clf_gender):
        self.network = network
        self.clf_age, self.clf_race, self.clf_gender = clf_age, clf_race, clf_gender

    @classmethod
    def from_pipeline(cls, networks, classifiers):
        """The last three network / classifier pairs, as ``estimate_age_race_gender`` indexes them
        (``num_networks - 3 / - 2 / - 1``; the race and gender entries reuse the age features)."""
        return cls(networks[len(networks) - 3], classifiers[-3], classifiers[-2], classifiers[-1])

    def age_patches(self, images, detections, age_subimage_width=96, age_subimage_height=96):
        """``age_subimages_arr`` of every detection, on the device: images = list of 2-D uint8 CUDA tensors (or numpy
        arrays), detections = list (one per image) of (M, 10) rows [box, angle, eye_l_x, eye_l_y, eye_r_x, eye_r_y, conf].
        Returns (TILED float32 CUDA tensor of contrast-normalised 96 x 96 patches, number of faces)."""
        import torch
        dev = torch.device("cuda", self.network.device)
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            sp = C.c_void_p(stream) if stream else None
            imgs = [im if torch.is_tensor(im) else torch.as_tensor(np.ascontiguousarray(im, dtype=np.uint8), device=dev) for im in images]
            params, index = [], []
            for k, (im, det) in enumerate(zip(imgs, detections)):
                det = np.asarray(det, dtype=np.float64).reshape(-1, 10)
                if len(det):
                    params.append(normalize.face_params(det[:, 5:9], int(im.shape[1]), int(im.shape[0])))
                    index += [k] * len(det)
            n = len(index)
            dim = age_subimage_width * age_subimage_height
            n_pad = (n + _lib.TILE - 1) // _lib.TILE * _lib.TILE
            patches = torch.zeros(max(n_pad, _lib.TILE) * dim, dtype=torch.float32, device=dev)
            if n == 0:
                return patches, 0
            xt, yt = normalize.age_tables(age_subimage_width, age_subimage_height)
            d_par = torch.as_tensor(np.concatenate(params), device=dev)
            d_idx = torch.as_tensor(np.asarray(index, dtype=np.int32), device=dev)
            d_ptr = torch.tensor([t.data_ptr() for t in imgs], dtype=torch.int64, device=dev)
            d_hw = torch.tensor([[t.shape[0], t.shape[1]] for t in imgs], dtype=torch.int32, device=dev)
            d_xt, d_yt = torch.as_tensor(xt, device=dev), torch.as_tensor(yt, device=dev)
            _lib.check(lib.hgsfa_age_crop_device(C.c_void_p(d_ptr.data_ptr()), C.c_void_p(d_hw.data_ptr()), C.c_void_p(d_idx.data_ptr()),
                                                 C.c_void_p(d_par.data_ptr()), n, C.c_void_p(d_xt.data_ptr()), C.c_void_p(d_yt.data_ptr()),
                                                 age_subimage_width, age_subimage_height, C.c_void_p(patches.data_ptr()), sp))
            _lib.check(lib.hgsfa_contrast_avg_std_device(C.c_void_p(patches.data_ptr()), n, dim, normalize.AGE_OBJ_AVG,
                                                         normalize.AGE_OBJ_STD, sp))
            torch.cuda.current_stream(dev).synchronize()       # the small parameter tensors above go out of scope
        return patches, n

    def estimate_detections(self, images, detections, long_text=True):
        """``estimate_age_race_gender`` for every detection of every image in ONE batch: crops, flow, three heads.
        Returns a list (one entry per image) of dicts with ``age``, ``age_std``, ``race``, ``gender`` (strings like
        the reference) and ``race_confidence``, ``gender_confidence``."""
        patches, n = self.age_patches(images, detections)
        counts = [len(np.asarray(d).reshape(-1, 10)) for d in detections]
        if n == 0:
            return [dict(age=np.zeros(0), age_std=np.zeros(0), race=[], gender=[], race_confidence=np.zeros(0),
                         gender_confidence=np.zeros(0)) for _ in counts]
        sl = self.network.execute_torch(patches, layout=_lib.TILED, n=n).cpu().numpy().astype(np.float64)
        c = self.clf_age
        age, age_std = c.regression(sl[:, 0:c.input_dim], c.avg_labels, estimate_std=True)
        c = self.clf_race
        race = c.regression(sl[:, 0:c.input_dim], c.avg_labels)
        c = self.clf_gender
        gender = c.regression(sl[:, 0:c.input_dim], c.avg_labels)
        races = map_real_race_labels_to_strings(race, long_text)
        genders = map_real_gender_labels_to_strings(gender, long_text)
        out, pos = [], 0
        for m in counts:
            sel = slice(pos, pos + m)
            out.append(dict(age=age[sel], age_std=age_std[sel], race=races[sel], gender=genders[sel],
                            race_confidence=np.abs(race[sel]) / 2.0, gender_confidence=np.abs(gender[sel])))
            pos += m
        return out

    def estimate(self, age_subimages_arr, estimate_age=True, estimate_race=True, estimate_gender=True, long_text=True):
        """Returns ``(age_estimates, age_stds, race_estimates, gender_estimates)`` like the reference, plus
        ``race_confidences`` and ``gender_confidences`` as a dict in fifth position.  Faces that are not estimated
        keep the reference's defaults (age 0, race / gender label 10 -> the mapping raises, as it does there)."""
        x = np.asarray(age_subimages_arr)
        n = x.shape[0]
        age = np.zeros(n)
        age_std = np.zeros(n)
        race = 10 * np.ones(n)
        gender = 10 * np.ones(n)
        if n:
            sl = self.network.execute(x)
            if estimate_age:
                c = self.clf_age
                age, age_std = c.regression(sl[:, 0:c.input_dim], c.avg_labels, estimate_std=True)
            if estimate_race:
                c = self.clf_race
                race = c.regression(sl[:, 0:c.input_dim], c.avg_labels)
            if estimate_gender:
                c = self.clf_gender
                gender = c.regression(sl[:, 0:c.input_dim], c.avg_labels)
        conf = dict(gender_confidences=np.abs(gender), race_confidences=np.abs(race) / 2.0)
        return (age, age_std, map_real_race_labels_to_strings(race, long_text),
                map_real_gender_labels_to_strings(gender, long_text), conf)
