"""Generate tests/golden/cascade_golden.json: the float64 oracle's cascade trace on BASELINE config 1
(sample_images/TNS-Group.jpg prescaled to 1000 x 750, smallest_face 0.1 -> 1 308 windows) with the synthetic U11L_64
model set (tests/cascade_models.py; the shipped flows are absent).  Pins the oracle loop + model set on this fixture:
tests/test_gpu_cascade.py compares the GPU cascade with the oracle run live AND with these committed numbers."""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cascade_models as cm  # noqa: E402
from oracle import cascade as ocascade  # noqa: E402

CUT = [0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.6]
m = cm.cached_models(spec="U11L_64")
img = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "tns_group_1000x750.png")))
purged, tr = ocascade.detect_image(img, m["header"], m["network_types"], m["networks"], m["classifiers"], 0.1,
                                   m["num_face_stages"], cut_offs_face=CUT, eye_header=m["header_eye"])
out = {"image": "tns_group_1000x750.png", "smallest_face": 0.1, "cut_offs_face": CUT, "models": "cascade_models.cached_models(spec='U11L_64', seed=0)",
       "stage_counts": [int(c) for c in tr["stage_counts"]], "raw": np.round(tr["raw"], 6).tolist(),
       "purged": np.round(purged, 6).tolist()}
with open(os.path.join(ROOT, "tests", "golden", "cascade_golden.json"), "w") as f:
    json.dump(out, f, indent=0)
print(out["stage_counts"], len(out["raw"]), len(out["purged"]))
