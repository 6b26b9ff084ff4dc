"""Pipeline description files and model loading.

Replaces ``load_networks_from_pipeline`` (reference ``face_analysis.py:374-493``): same text format
(line 1 count; line 2 ``net_Dx net_Dy net_Dang net_mins net_maxs sub_w sub_h reg_w reg_h``; line 3 eye
``Dx Dy mins maxs sub_w sub_h reg_w reg_h``; line 4 age likewise; then (type, network file, classifier
file) triples whose names lose their ``.pckl`` by ``[0:-5]``), same return tuple -- but every entry is a GPU
object, and a file referenced several times (FaceCentering2_1468510885 appears 4x) is loaded and
compiled once.
"""
from __future__ import annotations

import os

from .classifier import GpuGaussianClassifier
from .gpuflow import GpuFlow
from .pickles import load_obj

# FaceDetectUpdated.py:98,110-115,122
CUT_OFFS_FACE = [0.99, 0.95, 0.85, 0.8, 0.7, 0.6, 0.5, 0.45, 0.10, 0.05]
DEFAULTS = dict(patch_overlap_sampling=1.1, patch_overlap_posx_posy=1.1, tolerance_scale_deviation=1.1,
                tolerance_angle_deviation=1.1, tolerance_posxy_deviation=1.1, prescale_size=1000)


def parse_pipeline(path):
    with open(path, "r") as f:
        lines = [ln.rstrip("\n").rstrip("\r") for ln in f.readlines()]
    n = int(lines[0])
    h = lines[1].split(" ")
    net = (int(h[0]), int(h[1]), float(h[2]), float(h[3]), float(h[4]), int(h[5]), int(h[6]), int(h[7]), int(h[8]))
    e = lines[2].split(" ")
    eye = (int(e[0]), int(e[1]), float(e[2]), float(e[3]), int(e[4]), int(e[5]), int(e[6]), int(e[7]))
    a = lines[3].split(" ")
    age = (int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]), int(a[7]))
    types, nets, clfs = [], [], []
    for i in range(n):
        types.append(lines[4 + 3 * i].rstrip())
        nets.append(lines[5 + 3 * i].rstrip()[0:-5])
        clfs.append(lines[6 + 3 * i].rstrip()[0:-5])
    return dict(num_networks=n, net=net, eye=eye, age=age, network_types=types, network_filenames=nets,
                classifier_filenames=clfs)


def load_networks_from_pipeline(pipeline_filename, cache_obj=None, networks_base_dir="SavedNetworks",
                                classifiers_base_dir="SavedClassifiers", verbose_pipeline=True, verbose_networks=True,
                                device=0):
    """Same signature and return value as the reference function; ``cache_obj`` is accepted and unused."""
    p = parse_pipeline(pipeline_filename)
    network_types = p["network_types"] + ["None"] * (18 - len(p["network_types"]))
    flows = {}
    networks = []
    for name in p["network_filenames"]:
        if name == "None0":
            networks.append(None)
            continue
        if name not in flows:
            obj = load_obj(networks_base_dir, os.path.basename(name))
            if isinstance(obj, (list, tuple)):      # (flow, layers, benchmark, Network): keep only the flow
                obj = obj[0]
            flows[name] = GpuFlow(obj, device=device)
        networks.append(flows[name])
    heads = {}
    classifiers = []
    for name in p["classifier_filenames"]:
        if name not in heads:
            heads[name] = GpuGaussianClassifier(load_obj(classifiers_base_dir, os.path.basename(name)), device=device)
        classifiers.append(heads[name])
    return p["net"], p["eye"], p["age"], p["num_networks"], network_types, networks, classifiers
