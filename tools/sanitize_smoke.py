"""Smallest end-to-end case for compute-sanitizer: crop -> tiny flow (all pass shapes: clone layer, folded and
two-pass iGSFA, K-split) -> head -> controller -> compaction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, extract_subimages, synthetic  # noqa: E402
from conftest import load_classifiers  # noqa: E402

flow = synthetic.make_flow("tiny", seed=3)
x = synthetic.synthetic_patches(300, (16, 16), 1)
for mode in ("fold", "two_pass"):
    g = GpuFlow(flow, igsfa_mode=mode)
    y = g.execute(x)
    y2 = g.execute(x.astype(np.float32))
    assert np.isfinite(y).all() and np.allclose(y, y2, atol=1e-3)
    g.close()
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (100, 130), dtype=np.uint8)
boxes = np.array([[3.0, 4.0, 60.0, 61.0], [-5.0, -6.0, 40.0, 39.0], [80.0, 50.0, 140.0, 110.0]])
p = extract_subimages(img, boxes, np.array([0.0, 10.0, 0.0]), (64, 64), 0, np.uint8)
assert p.shape == (3, 4096)
clf = load_classifiers()[1]
h = GpuGaussianClassifier(clf)
r = h.regression(np.asarray(clf.means), clf.avg_labels)
assert r.shape == (10,)
import cascade_models as cm  # noqa: E402
from pyfaceanalysis_b200.cascade import FaceDetector  # noqa: E402
m = cm.cached_models()
flows, heads = {}, {}
nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f)) for f in m["networks"]]
clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c)) for c in m["classifiers"]]
det = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=[0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.1, 0.6],
                   header_eye=m["header_eye"])
out = det.detect([cm.test_scene(5)[0]], smallest_face=0.3)
print("sanitize smoke ok", y.shape, len(out[0]))
