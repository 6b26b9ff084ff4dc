"""Age / race / gender estimation on normalised face crops: the compute part of the reference's
``estimate_age_race_gender`` (``face_analysis.py:1170-1306``).

Reference, per final face::

    sl = networks[num_networks - 3].execute(age_subimages_arr)                       # (1, 9216) -> features
    age, age_std = classifiers[num_networks - 3].regression(sl[:, :D], avg_labels, estimate_std=True)
    race   = classifiers[num_networks - 2].regression(sl[:, :D], avg_labels)
    gender = classifiers[num_networks - 1].regression(sl[:, :D], avg_labels)
    gender_confidences = |gender|;  race_confidences = |race| / 2
    strings through map_real_gender_labels_to_strings / map_real_race_labels_to_strings

Here: one batched flow execute over all faces and three batched heads.  NOT built: the
96x96 crop itself -- ``normalize_image`` (``face_normalization_tools.py:111-329``) composes an integer EXTENT crop,
cuicuilco's ``rotate_improved(BICUBIC)`` and a BICUBIC EXTENT resample, followed by cuicuilco's
``load_image_data_monoprocessor`` sub-sampling with contrast enhancement; two of those four steps live in the
un-vendored cuicuilco.  The caller hands in the (N, 9216) patch matrix the reference calls ``age_subimages_arr``.
"""
from __future__ import annotations

import numpy as np


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
