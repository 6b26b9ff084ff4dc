"""Entry points that take raw device buffers run on the device that owns the buffers, whatever device the calling thread
has selected (hgsfa_cascade_update_device / compact / gather / crop / contrast launched on the current device)."""
import numpy as np
import pytest

import cascade_models as cm

pytestmark = pytest.mark.gpu


def test_detector_on_device_1_while_device_0_is_current():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier
    from pyfaceanalysis_b200.cascade import FaceDetector
    m = cm.cached_models()
    CUT = [0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.6]
    img = cm.test_scene(5)[0]
    outs = []
    for d in (0, 1):
        flows, heads = {}, {}
        nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f, device=d)) for f in m["networks"]]
        clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c, device=d)) for c in m["classifiers"]]
        det = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=CUT, header_eye=m["header_eye"], device=d)
        torch.cuda.set_device(0)                       # the calling thread stays on device 0
        outs.append(det.detect([img], smallest_face=0.2)[0])
        assert torch.cuda.current_device() == 0
    assert np.array_equal(outs[0], outs[1])
