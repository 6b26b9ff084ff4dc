#!/usr/bin/env python
"""BASELINE.json configs[3]-style run: eye-position and age / race / gender flows on 100k synthetic face crops.

The shipped flows are stripped from the reference; the stand-ins are the synthetic networks of
pyfaceanalysis_b200/synthetic.py with the shapes the reference's pipelines use: a 64x64 eye network (U11L_64: the
shipped eye classifiers are named "...Ultra Thin 11 Layer Network...REyePosXY..."; reference eye patches are 64x64,
FaceDetectUpdated.py:149-151) followed by the two REAL shipped eye heads (EyeLX 12 x 50, EyeLY 10 x 50), and the 96x96
age network (U11L_96; face_analysis.py:1170-1181), float32 contrast-normalised crops, three Gaussian heads on the age
features (real shipped parameters: Age / RaceC / GenderC of tests/golden/classifiers.npz).

    python tools/bench_flows.py [--n 100000] [--steps 5] [--warmup 2]

Prints one JSON line per flow: crops/s with the crops resident in HBM (tiled float32, as
hgsfa_crop_extent_device + hgsfa_contrast_avg_std_device leave them), heads included for the age flow.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import ctypes as C
    import types
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, _lib, synthetic

    dev = torch.device("cuda:0")
    lib = _lib.load()
    heads, eye_heads = [], []      # the real shipped "Generalize" heads (age 4x39, race 5x2, gender 5x2) and eye heads
    z = np.load(os.path.join(ROOT, "tests", "golden", "classifiers.npz"))
    for k, name in enumerate(z["names"]):
        is_eye = "REyePosXY" in str(name) or "EyeL" in str(name)
        if "Generalize" not in str(name) and not is_eye:
            continue
        key = "c%02d" % k
        clf = types.SimpleNamespace(means=list(z[key + "_means"]), inv_covs=list(z[key + "_inv_covs"]),
                                    _sqrt_def_covs=list(z[key + "_sqrt_def_covs"]), p=list(z[key + "_p"]),
                                    labels=list(z[key + "_labels"]), avg_labels=z[key + "_avg_labels"])
        clf._input_dim = clf.input_dim = clf.means[0].shape[0]
        h = GpuGaussianClassifier(clf)
        (eye_heads if is_eye else heads).append((str(name), h, torch.as_tensor(np.asarray(clf.avg_labels, dtype=np.float64), device=dev)))
    eye_heads = eye_heads[:2]
    for spec, side, with_heads in (("U11L_64", 64, eye_heads), ("U11L_96", 96, heads)):
        flow = synthetic.cached_flow(spec)
        g = GpuFlow(flow)
        n = args.n
        n_pad = (n + _lib.TILE - 1) // _lib.TILE * _lib.TILE
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        # contrast-normalised crops: zero mean, std 0.16 (face_analysis.py:1184-1195), tiled layout
        x = torch.randn(n_pad * side * side, device=dev, generator=gen, dtype=torch.float32) * 0.16

        reg = torch.empty(n, dtype=torch.float64, device=dev)

        def step():
            sl = g.execute_torch(x, layout=_lib.TILED, n=n)
            if with_heads:
                for _, h, labels in with_heads:
                    _lib.check(lib.hgsfa_gauss_regress_device(h.handle, C.c_void_p(sl.data_ptr()), _lib.F32, n, sl.stride(0),
                                                              C.c_void_p(labels.data_ptr()), C.c_void_p(reg.data_ptr()),
                                                              None, None, None, None))
            return sl

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        fl = g.flops(n, np.float32)
        print(json.dumps({"metric": "flow crops/sec", "flow": spec, "input": "%dx%d float32 (tiled, resident)" % (side, side),
                          "n": n, "value": n / (ms * 1e-3), "unit": "crops/s", "ms_per_step": ms,
                          "algorithmic_tflops": fl["algorithmic"] / (ms * 1e-3) / 1e12,
                          "heads": [h[0] for h in with_heads],
                          "engines": sorted({op.engine for op in g.spec.ops})}))
        g.close()


if __name__ == "__main__":
    main()
