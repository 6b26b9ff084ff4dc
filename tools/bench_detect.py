"""Secondary benchmark (BASELINE.json configs[2]): full detection pyramid, images/s.

64 synthetic 1000x562 'L' images (= 1920x1080 after the reference's default prescale to <= 1000 px,
FaceDetectUpdated.py:551-559), smallest_face = 0.05 -> 7 452 windows per image, 476 928 per batch; the 17 face
stages + eye stage of Pipelines/Pipeline_experimental.txt with the synthetic U11L_64 model set
(tests/cascade_models.py; the shipped flows were stripped).  GPU: one batched FaceDetector.detect() per step,
images already decoded in host memory (upload, grid, cascade, eye stage, purge inside the timed region).
CPU baseline: the oracle's restatement of the reference loop on a bounded sample of images.
Prints one JSON object.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-images", type=int, default=1)
    ap.add_argument("--spec", default="U11L_64")
    ap.add_argument("--smallest-face", type=float, default=0.05)
    args = ap.parse_args()
    import torch
    import cascade_models as cm
    from oracle import cascade as ocascade
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier
    from pyfaceanalysis_b200.cascade import FaceDetector
    from threadpoolctl import threadpool_limits

    m = cm.cached_models(spec=args.spec)
    flows, heads = {}, {}
    nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f)) for f in m["networks"]]
    clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c)) for c in m["classifiers"]]
    rng = np.random.default_rng(0)
    images = []
    for k in range(args.images):
        faces = [(rng.uniform(80, 920), rng.uniform(80, 480), rng.uniform(40, 160), rng.uniform(-10, 10)) for _ in range(4)]
        images.append(cm.render_scene(562, 1000, faces, 1000 + k))
    # The synthetic heads are not trained to the reference's operating point: calibrate the Disc cut-offs on one
    # image so that the funnel has the shape a cascade is built for (stage 0 sees every window, few survive).
    keep = {"Disc1": 0.04, "Disc3": 0.4, "Disc5": 0.5, "Disc7": 0.6, "Disc9": 0.5}
    cut = [1e30] * 10
    for name, frac in keep.items():
        d0 = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=cut, header_eye=None)
        _, tr0 = d0.detect(images[:1], smallest_face=args.smallest_face, return_trace=True)
        sc = tr0["disc_scores"].get(name)
        cut[int(name[-1])] = float(np.nanquantile(sc, frac)) if sc is not None and len(sc) else 0.0
    det = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=cut, header_eye=m["header_eye"])
    for _ in range(args.warmup):
        out, tr = det.detect(images, smallest_face=args.smallest_face, return_trace=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, tr = det.detect(images, smallest_face=args.smallest_face, return_trace=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    res = {"metric": "detect images/sec", "value": args.images / dt, "unit": "images/s", "ms_per_batch": dt * 1e3,
           "images": args.images, "image_size": [1000, 562], "smallest_face": args.smallest_face,
           "windows_per_batch": int(tr["n_windows"]), "stage_counts": [int(c) for c in tr["stage_counts"]],
           "detections": int(sum(len(o) for o in out)), "calibrated_cut_offs": cut, "models": "synthetic " + args.spec + " cascade (tests/cascade_models.py)"}
    if os.environ.get("HGSFA_DETECT_PROFILE"):
        det.detect(images, smallest_face=args.smallest_face)
        res["phase_ms"] = {k: round(v * 1e3, 2) for k, v in det.last_profile.items()}
        print(json.dumps(res["phase_ms"]), file=sys.stderr)
    if args.cpu_images <= 0:
        print(json.dumps(res))
        return
    threads = min(12, os.cpu_count() or 1)
    with threadpool_limits(limits=threads):
        t0 = time.perf_counter()
        for img in images[:args.cpu_images]:
            ocascade.detect_image(img, m["header"], m["network_types"], m["networks"], m["classifiers"], args.smallest_face,
                                  m["num_face_stages"], cut_offs_face=cut, eye_header=m["header_eye"])
        cdt = (time.perf_counter() - t0) / args.cpu_images
    res["cpu_baseline"] = {"value": 1.0 / cdt, "unit": "images/s", "cores": threads, "kind": "port",
                           "sample": "%d image(s), float64 numpy oracle of the reference loop" % args.cpu_images}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
