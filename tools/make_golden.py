"""Generate the committed fixtures under tests/golden/ from the reference tree (run in the build
container, where /root/reference exists; the GPU box only ever sees the fixtures).

  classifiers.npz : parameters of the 19 shipped SavedClassifiers/*.pckl (means, inv_covs,
                    _sqrt_def_covs, p, labels, avg_labels) -- model data, float64, unmodified
  pipeline.json   : Pipelines/Pipeline_experimental.txt parsed like face_analysis.py:374-443 plus the
                    controller constants of FaceDetectUpdated.py:98,110-115
  crop_golden.npz : Pillow 12.2 Image.transform(EXTENT, NEAREST|BILINEAR) outputs on a seeded synthetic
                    image for random, overhanging and adversarial boxes
  grid_golden.json: window counts per scale for the BASELINE.md image sizes, from the oracle grid
                    (which is the reference arithmetic of face_analysis.py:575-669)
  tns_group_1000x750.png : BASELINE config 1 input -- sample_images/TNS-Group.jpg (README.md:43) opened in mode 'L' and
                    prescaled exactly like FaceDetectUpdated.py:551-559 (Pillow resize NEAREST to 1000 x 750)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from pyfaceanalysis_b200 import pickles  # noqa: E402
from oracle import grid as ogrid  # noqa: E402


def classifiers():
    d = os.path.join(REF, "SavedClassifiers")
    arrays = {}
    names = sorted(f for f in os.listdir(d) if f.endswith(".pckl"))
    for k, fn in enumerate(names):
        clf = pickles.load_obj(d, fn[:-5])
        key = "c%02d" % k
        arrays[key + "_means"] = np.asarray([np.asarray(m, dtype=np.float64) for m in clf.means])
        arrays[key + "_inv_covs"] = np.asarray([np.asarray(m, dtype=np.float64) for m in clf.inv_covs])
        arrays[key + "_sqrt_def_covs"] = np.asarray([float(v) for v in clf._sqrt_def_covs])
        arrays[key + "_p"] = np.asarray([float(v) for v in clf.p])
        arrays[key + "_labels"] = np.asarray([float(v) for v in clf.labels])
        arrays[key + "_avg_labels"] = np.asarray(clf.avg_labels, dtype=np.float64)
    arrays["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(OUT, "classifiers.npz"), **arrays)
    return names


def pipeline():
    path = os.path.join(REF, "Pipelines", "Pipeline_experimental.txt")
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f.readlines()]
    n = int(lines[0])
    h = lines[1].split(" ")
    net = [int(h[0]), int(h[1]), float(h[2]), float(h[3]), float(h[4]), int(h[5]), int(h[6]), int(h[7]), int(h[8])]
    e = lines[2].split(" ")
    eye = [int(e[0]), int(e[1]), float(e[2]), float(e[3]), int(e[4]), int(e[5]), int(e[6]), int(e[7])]
    a = lines[3].split(" ")
    age = [int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]), int(a[7])]
    types, nets, clfs = [], [], []
    for i in range(n):
        types.append(lines[4 + 3 * i].rstrip())
        nets.append(lines[5 + 3 * i].rstrip()[0:-5])
        clfs.append(lines[6 + 3 * i].rstrip()[0:-5])
    out = dict(num_networks=n, net=net, eye=eye, age=age, network_types=types, network_filenames=nets,
               classifier_filenames=clfs,
               cut_offs_face=[0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.05],
               patch_overlap_sampling=1.1, patch_overlap_posx_posy=1.1, tolerance_scale_deviation=1.1,
               tolerance_angle_deviation=1.1, tolerance_posxy_deviation=1.1, prescale_size=1000)
    with open(os.path.join(OUT, "pipeline.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


def crop_golden():
    from PIL import Image
    rng = np.random.default_rng(20180329)
    H, W = 240, 320
    yy, xx = np.mgrid[0:H, 0:W]
    img = (128 + 60 * np.sin(xx / 9.0) * np.cos(yy / 13.0) + 40 * rng.standard_normal((H, W))).clip(0, 255).astype(np.uint8)
    pim = Image.fromarray(img, "L")
    boxes = []
    for _ in range(40):
        s = rng.uniform(12, 300)
        x0 = rng.uniform(-50, W - 5)
        y0 = rng.uniform(-50, H - 5)
        boxes.append((x0, y0, x0 + s - 1, y0 + s - 1))
    for x0 in (0.9, 0.7, 1.3, 2.1, 0.3, 10.9):          # coordinates that land on integers
        for a in (0.2, 0.6, 1.4, 0.3, 2.2, 0.7):
            boxes.append((x0, x0, x0 + 64 * a, x0 + 64 * a))
    boxes = np.asarray(boxes, dtype=np.float64)
    near = np.stack([np.asarray(pim.transform((64, 64), Image.EXTENT, tuple(b), Image.NEAREST)) for b in boxes])
    bil = np.stack([np.asarray(pim.transform((64, 64), Image.EXTENT, tuple(b), Image.BILINEAR)) for b in boxes])
    np.savez_compressed(os.path.join(OUT, "crop_golden.npz"), image=img, boxes=boxes, nearest=near, bilinear=bil)


def grid_golden(pipe):
    cases = {}
    for name, (w, h, sf, prescale) in {
        "tns_group_0.1": (3648, 2736, 0.1, True),
        "fhd_0.05_prescaled": (1920, 1080, 0.05, True),
        "fhd_0.05": (1920, 1080, 0.05, False),
        "uhd_0.02_prescaled": (3840, 2160, 0.02, True),
        "uhd_0.02": (3840, 2160, 0.02, False),
    }.items():
        if prescale:
            w2, h2, _ = ogrid.prescaled_size(w, h)
        else:
            w2, h2 = w, h
        wins = ogrid.enumerate_windows(w2, h2, pipe["net"], sf)
        cases[name] = dict(width=w2, height=h2, smallest_face=sf, counts=[int(len(c)) for _, c, _ in wins],
                           sampling_values=[float(s).hex() for s, _, _ in wins],
                           first_box=[float(v).hex() for v in wins[0][1][1]],
                           last_box=[float(v).hex() for v in wins[-1][1][-1]])
    with open(os.path.join(OUT, "grid_golden.json"), "w") as f:
        json.dump(cases, f, indent=1)
    return cases


def tns_image():
    from PIL import Image
    im = Image.open(os.path.join(REF, "sample_images", "TNS-Group.jpg")).convert("L")
    factor = max(im.size[0] * 1.0 / 1000, im.size[1] * 1.0 / 1000)
    size = (int(im.size[0] / factor), int(im.size[1] / factor))
    small = im.resize(size, Image.NEAREST)
    small.save(os.path.join(OUT, "tns_group_1000x750.png"), optimize=True)
    return im.size, small.size


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    print("TNS-Group:", tns_image())
    names = classifiers()
    print("classifiers:", len(names))
    pipe = pipeline()
    print("pipeline:", pipe["num_networks"], pipe["net"])
    crop_golden()
    g = grid_golden(pipe)
    for k, v in g.items():
        print(k, v["width"], v["height"], sum(v["counts"]), v["counts"])
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
