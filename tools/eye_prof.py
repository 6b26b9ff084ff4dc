"""Builder tool: host wall-clock of the pieces of the eye stage of FaceDetector.detect (synchronising timers) on the bench's
configs[2] batch."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import cascade_models as cm  # noqa: E402
from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, _lib  # noqa: E402
from pyfaceanalysis_b200 import cascade as casc  # noqa: E402

dev = torch.device("cuda:0")
cfg = bench.DETECT_CONFIGS[2]
m = cm.cached_models(spec=bench.FLOW_SPEC)
flows, heads = {}, {}
nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f, device=0)) for f in m["networks"]]
clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c, device=0)) for c in m["classifiers"]]
base = bench._detect_scenes(cm, cfg, 8)
host = [torch.from_numpy(np.ascontiguousarray(np.roll(base[k % 8], (k * 37) % cfg["hw"][1], axis=1))) for k in range(64)]
cut = [float(c) for c in os.environ.get("CUTS", "").split(",")] if os.environ.get("CUTS") else None
if cut is None:
    keep = {"Disc1": 0.04, "Disc3": 0.4, "Disc5": 0.5, "Disc7": 0.6, "Disc9": 0.5}
    cut = [1e30] * 10
    cal = casc.FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=cut, header_eye=None, device=0)
    cal_img = cal.prescale([host[0].to(dev)])[0]
    for name, frac in keep.items():
        cal.cut_offs = list(cut)
        _, tr0 = cal.detect([cal_img], smallest_face=cfg["smallest_face"], return_trace=True)
        sc = tr0["disc_scores"].get(name)
        cut[int(name[-1])] = float(np.nanquantile(sc, frac)) if sc is not None and len(sc) else 0.0
det = casc.FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=cut, header_eye=m["header_eye"], device=0)
imgs = det.prescale([h.to(dev) for h in host])

T = {}


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w


det._find_eyes = timed("find_eyes total", det._find_eyes)
det.eye_net.execute_torch = timed("  eye flow", det.eye_net.execute_torch)
det._regress_orig = det._regress
casc.purge_detections_orig = casc.purge_detections
casc.purge_detections = timed("purge (all images)", casc.purge_detections)
casc.approximate_eye_boxes = timed("approximate_eye_boxes", casc.approximate_eye_boxes)
casc.group_confidences = timed("group_confidences", casc.group_confidences)
lib = _lib.load()
for fname in ("hgsfa_crop_extent_batch_device", "hgsfa_contrast_avg_std_device", "hgsfa_gauss_regress_device",
              "hgsfa_compact_index_device", "hgsfa_gather_rows_device", "hgsfa_cascade_update_device"):
    setattr(lib, fname, timed("  lib." + fname, getattr(lib, fname)))
for _ in range(2):
    det.detect(imgs, smallest_face=cfg["smallest_face"])
T.clear()
os.environ["HGSFA_DETECT_PROFILE"] = "1"
reps = 3
for _ in range(reps):
    det.detect(imgs, smallest_face=cfg["smallest_face"])
print({k: round(v / reps * 1e3, 3) for k, v in T.items()})
print({k: round(v * 1e3, 3) for k, v in det.last_profile.items()})
