"""GPU parity of the age stage: ``normalize_image`` + 96 x 96 sub-sampling as one kernel (csrc/crop.cu: age_crop_kernel)
against the oracle's restatement (oracle/normalize.py, pinned against Pillow in tests/test_oracle_normalize.py), and
``estimate_age_race_gender`` wired into the detector (face_analysis.py:1170-1306, FaceDetectUpdated.py:1187)."""
import numpy as np
import pytest

import cascade_models as cm
from oracle import crop as ocrop
from oracle import gauss as ogauss
from oracle import nodes as onodes
from oracle import normalize as onorm

pytestmark = pytest.mark.gpu


def _untile(t, n, dim):
    t = t.cpu().numpy().reshape(-1, dim, 128)
    return np.concatenate([t[k].T for k in range(t.shape[0])])[:n]


def _estimator(u11l96_flow, classifiers):
    from pyfaceanalysis_b200 import AttributeEstimator, GpuFlow, GpuGaussianClassifier
    heads = {("Age" if "Age" in c.name else "Race" if "Race" in c.name else "Gender"): c
             for c in classifiers if "Generalize" in c.name}
    g = GpuFlow(u11l96_flow)
    return AttributeEstimator(g, *[GpuGaussianClassifier(heads[k]) for k in ("Age", "Race", "Gender")]), g, heads


def test_age_crop_bit_exact(u11l96_flow, classifiers):
    """Every sample of the 96 x 96 patch goes through crop -> rotate(BICUBIC) -> EXTENT(BICUBIC) -> NEAREST exactly as
    Pillow would compute the three uint8 images: the un-normalised patches equal the oracle's bytes, for upright,
    tilted, tiny, border-overhanging faces and the rotate(0) = copy shortcut."""
    import ctypes as C
    import torch
    from pyfaceanalysis_b200 import _lib, normalize
    est, g, _ = _estimator(u11l96_flow, classifiers)
    imgs = [cm.test_scene(21, 300, 400, 3)[0], cm.test_scene(22, 240, 320, 2)[0]]
    eyes = [np.array([[120.3, 90.2, 171.8, 96.9], [60.0, 100.0, 100.0, 100.0], [250.5, 140.25, 330.75, 112.5],
                      [5.0, 20.0, 60.0, 35.0], [300.0, 260.0, 392.0, 268.0]]),
            np.array([[100.0, 60.0, 124.0, 58.0], [10.0, 200.0, 150.0, 230.0]])]
    dets = [np.concatenate([np.zeros((len(e), 5)), e, np.full((len(e), 1), 0.3)], axis=1) for e in eyes]
    patches, n = est.age_patches(imgs, dets)
    assert n == 7
    got = _untile(patches, n, 9216)
    ref = np.concatenate([onorm.age_subimages(im, d) for im, d in zip(imgs, dets)])
    assert np.allclose(got, ref, rtol=0, atol=2e-6)                       # contrast-normalised float32 vs float64
    # the bytes before contrast normalisation: call the kernel alone
    dev = torch.device("cuda", 0)
    t_imgs = [torch.as_tensor(im, device=dev) for im in imgs]
    par = np.concatenate([normalize.face_params(d[:, 5:9], im.shape[1], im.shape[0]) for im, d in zip(imgs, dets)])
    xt, yt = normalize.age_tables()
    raw = torch.zeros(128 * 9216, dtype=torch.float32, device=dev)
    args = [torch.tensor([t.data_ptr() for t in t_imgs], dtype=torch.int64, device=dev),
            torch.tensor([[t.shape[0], t.shape[1]] for t in t_imgs], dtype=torch.int32, device=dev),
            torch.as_tensor(np.array([0] * 5 + [1] * 2, dtype=np.int32), device=dev), torch.as_tensor(par, device=dev),
            torch.as_tensor(xt, device=dev), torch.as_tensor(yt, device=dev)]
    _lib.check(_lib.load().hgsfa_age_crop_device(*[C.c_void_p(a.data_ptr()) for a in args[:4]], 7,
                                                 C.c_void_p(args[4].data_ptr()), C.c_void_p(args[5].data_ptr()), 96, 96,
                                                 C.c_void_p(raw.data_ptr()), None))
    got_raw = _untile(raw, 7, 9216)
    box = np.asarray(onorm.age_box())[None, :]
    k = 0
    for im, d in zip(imgs, dets):
        for row in d:
            im2 = onorm.normalize_image(im, row[5:9])
            assert np.array_equal(got_raw[k], ocrop.extract_subimages(im2, box, None, (96, 96))[0]), k
            k += 1
    g.close()


def test_attribute_stage_in_detector(u11l96_flow, classifiers):
    """detect(..., estimate_attributes=True): the purged detections go through the age crop, the age flow and the three
    REAL shipped heads; equals the oracle evaluated the reference's way, per face, on the oracle's own crops."""
    from pyfaceanalysis_b200.cascade import FaceDetector, format_detections
    est, g, heads = _estimator(u11l96_flow, classifiers)
    m = cm.cached_models()
    from test_gpu_cascade import CUT, _gpu_models
    nets, clfs = _gpu_models(m)
    det = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=CUT, header_eye=m["header_eye"], attributes=est)
    images = [cm.test_scene(seed)[0] for seed in (5, 6)]
    got, attrs = det.detect(images, smallest_face=0.2, estimate_attributes=True)
    assert len(attrs) == 2 and sum(len(d) for d in got) > 0
    for im, d, a in zip(images, got, attrs):
        assert len(a["age"]) == len(d) == len(a["race"]) == len(a["gender"])
        if not len(d):
            continue
        sl = onodes.flow_execute(u11l96_flow, onorm.age_subimages(im, d))
        age, std = ogauss.regression(heads["Age"], sl[:, :4], heads["Age"].avg_labels, estimate_std=True)
        ok = ~np.isnan(age)
        # features within 1e-3 x std of the float64 flow; the heads are smooth in the features at that scale
        assert np.allclose(a["age"][ok], age[ok], rtol=0.05, atol=0.5), (a["age"], age)
        assert np.array_equal(np.isnan(a["age"]), np.isnan(age))
        text = format_detections(d, attributes=(a["age"], a["race"], a["gender"]))
        assert text.count("\n") == len(d) and (", Male, " in text or ", Female, " in text)
    with pytest.raises(ValueError):
        FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=CUT).detect(images, estimate_attributes=True)
    g.close()
