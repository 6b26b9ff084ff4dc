"""Batch mode (pyfaceanalysis_b200/batch.py; reference face_analysis.py:224-243, FaceDetectUpdated.py:533-559, 1258-1278)."""
import os

import numpy as np
import pytest

from pyfaceanalysis_b200 import batch
from pyfaceanalysis_b200.cascade import format_detections


def _write_images(tmp_path, n):
    from PIL import Image
    rng = np.random.default_rng(9)
    names, arrays = [], []
    for k in range(n):
        a = rng.integers(0, 256, (40 + 3 * k, 50 + k), dtype=np.uint8)
        name = str(tmp_path / ("im%d.png" % k))
        Image.fromarray(a, "L").save(name)
        names.append(name)
        arrays.append(a)
    return names, arrays


def test_read_batch_file_pairs_and_trailing_line(tmp_path):
    p = tmp_path / "batch.txt"
    p.write_text("a.jpg\nout_a.txt  \nb.jpg\r\nout_b.txt\nunpaired.jpg\n")
    images, outputs = batch.read_batch_file(str(p))
    assert images == ["a.jpg", "b.jpg"] and outputs == ["out_a.txt", "out_b.txt"]       # rstrip; the odd last line is dropped
    p.write_text("")
    assert batch.read_batch_file(str(p)) == ([], [])


def test_load_images_is_pillow_mode_l(tmp_path):
    from PIL import Image
    names, arrays = _write_images(tmp_path, 2)
    rgb = np.random.default_rng(1).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    Image.fromarray(rgb, "RGB").save(str(tmp_path / "c.png"))
    got = batch.load_images(names + [str(tmp_path / "c.png")])
    assert all(g.dtype == np.uint8 and g.flags["C_CONTIGUOUS"] for g in got)
    assert np.array_equal(got[0], arrays[0]) and np.array_equal(got[1], arrays[1])
    assert np.array_equal(got[2], np.asarray(Image.fromarray(rgb, "RGB").convert("L")))  # the reference's conversion
    assert batch.load_images([str(tmp_path / "c.png")], "RGB")[0].shape == (20, 30, 3)


class _FakeDetector:
    """Records what run_batch asks for; a detection per image derived from the image it was given."""

    def __init__(self):
        self.calls = []

    def prescale(self, images, prescale_size=None):
        self.calls.append(("prescale", len(images), prescale_size))
        return images

    def detect(self, images, smallest_face=0.2, benchmark=None, estimate_attributes=False):
        self.calls.append(("detect", len(images), smallest_face, estimate_attributes))
        dets = []
        for im in images:
            h, w = im.shape
            d = np.array([[1.4, 2.6, w - 0.5, h + 0.49, -3.25, 10, 11, 20, 21.5, 0.125],
                          [0, 0, w / 2.0, h / 2.0, 7.0, 3, 4, 5, 6, float(im[0, 0]) / 255.0]])
            dets.append(d[: 1 + (h % 2)])
        if estimate_attributes:
            return dets, [dict(age=np.full(len(d), 30.5), race=["White"] * len(d), gender=["Male"] * len(d)) for d in dets]
        return dets


@pytest.mark.parametrize("group", [1, 2, 64])
def test_run_batch_writes_the_reference_lines(tmp_path, group):
    names, arrays = _write_images(tmp_path, 5)
    outs = [str(tmp_path / ("out%d.txt" % k)) for k in range(5)]
    with open(outs[1], "w") as f:
        f.write("earlier run\n")                                                       # results are appended (mode 'a')
    det = _FakeDetector()
    res = batch.run_batch(det, (names, outs), smallest_face=0.1, group=group, decode_threads=3)
    want = _FakeDetector().detect(arrays)
    assert len(res) == 5 and all(np.array_equal(a, b) for a, b in zip(res, want))
    for k in range(5):
        text = open(outs[k]).read()
        assert text == ("earlier run\n" if k == 1 else "") + format_detections(want[k])
    detect_calls = [c for c in det.calls if c[0] == "detect"]
    assert [c[1] for c in detect_calls] == [min(group, 5 - s) for s in range(0, 5, group)]
    assert all(c[2] == 0.1 for c in detect_calls) and ("prescale", min(group, 5), 1000) in det.calls
    # no prescale, attributes written, right screen eye first, batch file on disk
    bf = tmp_path / "batch.txt"
    outs2 = [str(tmp_path / ("attr%d.txt" % k)) for k in range(5)]
    bf.write_text("".join("%s\n%s\n" % (a, b) for a, b in zip(names, outs2)))
    det2 = _FakeDetector()
    batch.run_batch(det2, str(bf), image_prescaling=False, group=group, estimate_attributes=True, right_screen_eye_first=True)
    assert not any(c[0] == "prescale" for c in det2.calls)
    for k in range(5):
        nk = len(want[k])
        assert open(outs2[k]).read() == format_detections(want[k], True, (np.full(nk, 30.5), ["White"] * nk, ["Male"] * nk))
    assert batch.run_batch(det2, ([], [])) == []
    with pytest.raises(ValueError):
        batch.run_batch(det2, (names, outs[:2]))


@pytest.mark.gpu
def test_run_batch_on_the_device_cascade(tmp_path):
    """Files -> Pillow decode -> device prescale -> cascade -> result files, against detect() on the same pixels."""
    from PIL import Image
    import cascade_models as cm
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier
    from pyfaceanalysis_b200.cascade import FaceDetector
    m = cm.cached_models()
    flows, heads = {}, {}
    nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f)) for f in m["networks"]]
    clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c)) for c in m["classifiers"]]
    from test_gpu_cascade import CUT
    det = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=CUT, header_eye=m["header_eye"])
    names, outs, arrays = [], [], []
    for k, seed in enumerate((5, 6, 8)):
        img = cm.test_scene(seed)[0]
        if k == 2:                                              # one image large enough to be prescaled (> 1000 px wide)
            img = np.ascontiguousarray(np.kron(img, np.ones((4, 4), dtype=np.uint8)))
        name = str(tmp_path / ("scene%d.png" % k))
        Image.fromarray(img, "L").save(name)
        names.append(name)
        outs.append(str(tmp_path / ("scene%d.txt" % k)))
        arrays.append(img)
    res = batch.run_batch(det, (names, outs), smallest_face=0.2, group=2)
    assert max(arrays[2].shape) > 1000
    for k in range(3):
        one = det.detect(det.prescale([arrays[k]]), smallest_face=0.2)[0]
        assert np.array_equal(res[k], one)
        assert open(outs[k]).read() == format_detections(one)
    assert sum(len(r) for r in res) > 0
