#!/usr/bin/env python
"""Headline benchmark: HiGSFA flow windows/s (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): forward pass of the FaceCentering2-shaped flow -- the synthetic
11-layer "ultra thin" HiGSFA network U11L_64 (pyfaceanalysis_b200/synthetic.py; the shipped flow pickles
were stripped from the reference) -- over a batch of 1 048 576 windows of 64x64 = 4096 uint8 pixels,
uniform integers 0..255, seed 12345600 (FaceDetectUpdated.py:146).  One step = one such batch per GPU.

  value : windows/s with the (N, 4096) uint8 window matrix already resident in HBM (row-major, as the
          reference hands it over); timed with CUDA events on the launching stream, max over ranks
  e2e   : the same through the public drop-in call GpuFlow.execute(x) with x in pinned HOST memory:
          host->device copy of the windows and device->host copy of the (N, 60) float64 features are
          inside the timed region
  roofline : dominant kernel of the step: hgsfa::front_kernel (layers 0-2 fused, tcgen05 kind::f16,
          2-piece FP16 split: ceiling = measured bf16 peak / 3), hgsfa::layer_tc_kernel (FP16 pieces: bf16 peak / 3; 3xTF32: / 6)
          or hgsfa::layer_kernel (packed FP32 FMA: measured FFMA2 peak); achieved = algorithmic flops of that kernel's
          ops / its CUDA-event time inside the timed steps; every kernel of the step is listed under "kernels"
  detect : BASELINE configs[2] on the same clock -- 64 synthetic 1920x1080 images per GPU, smallest_face 0.05,
          images/s with upload, device prescale, 17 face stages + eye stage, host purge and the gather of the
          detection lists inside the timed region; images are sharded over the ranks (pyfaceanalysis_b200/shard.py)
  cpu_baseline : the float64 numpy oracle (the reference cannot run: Python 2 + un-vendored mdp /
          cuicuilco) on a bounded sample on the box's host cores, BLAS threads = min(12, nproc)
          (FaceDetectUpdated.py:74)

Multi-GPU: windows are independent -> every rank processes its own batch (weak scaling), no collective
on the data path; torch.distributed is used only for the barrier and the max-over-ranks of the time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WINDOWS = 1 << 20
SEED = 12345600
FLOW_SPEC = "U11L_64"
METRIC = "HiGSFA flow windows/sec"
UNIT = "windows/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        p["_source"] = "MEASURED_PEAKS.json"
        return p
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "_source": "fallback (B200_PROFILING.md)"}


def _fp32_peak():
    """Measured FFMA peak of this pool's B200 (tools/microbench.cu, profiles/microbench_r01.json)."""
    path = os.path.join(ROOT, "profiles", "microbench_r01.json")
    try:
        with open(path) as f:
            m = json.load(f)
        return float(m["ffma2_tflops"]), "profiles/microbench_r01.json (fma.rn.f32x2, measured on this pool)"
    except Exception:
        return 74.4, "nominal 148 SM x 128 FMA x 2 x 1.965 GHz"


def _traffic_per_window():
    """DRAM bytes per window of the layer kernels, from the committed ncu launch list (profiles/)."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return float(t["layer_dram_bytes_per_window"]), t["source"], float(t["layer_share_of_step"])
        except Exception:
            continue
    return None, None, None


def _roofline(op_stats, steps, n, fl, kernel_ms_last, peaks, fp32_peak, fp32_src, traffic, tsrc, lshare):
    """Roofline of the dominant kernel of a step, plus every kernel of the step under "kernels" (DESIGN.md section 6):
    hgsfa::front_kernel (layers 0-2 in one launch, 2-piece FP16 split on tcgen05 kind::f16: three bf16-rate MMAs per
    algorithmic block -> ceiling = measured bf16 peak / 3), hgsfa::layer_tc_kernel (3xTF32 at half the bf16 rate:
    peak / 6) and hgsfa::layer_kernel (packed FP32 FMA).  Times are CUDA-event totals per op over the timed steps, on
    the launching stream; flops are the algorithmic ones of SURVEY.md 8d (hgsfa_plan_flops)."""
    eng = {}
    for st in op_stats:
        e = eng.setdefault(st["engine"], dict(ms=0.0, alg=0.0, exe=0.0, ops=0))
        e["ms"] += st["ms"] / steps
        e["alg"] += st["alg_flops"] * n
        e["exe"] += st["exe_flops"] * n
        e["ops"] += 1
    for e in eng.values():
        e["achieved"] = e["alg"] / (e["ms"] * 1e-3) / 1e12 if e["ms"] > 0 else None
    bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    bf16_txt = "%s bf16 %s %.1f TFLOP/s" % (peaks["_source"], "sustained" if "bf16_tflops_sustained" in peaks else "burst", bf16)
    ceilings = {
        "front": ("hgsfa::front_kernel (layers 0-2 fused, lane-resident; tcgen05 kind::f16)", "tensor", bf16 / 3.0,
                  bf16_txt + " / 3 (2-piece FP16 split = 3 MMAs per algorithmic block)"),
        "f16": ("hgsfa::layer_tc_kernel<F16> (one layer per launch; tcgen05 kind::f16 on 2-piece FP16 operands; hgsfa::back_kernel with "
                "HGSFA_BACK=1)", "tensor", bf16 / 3.0,
                bf16_txt + " / 3 (2-piece FP16 split = 3 MMAs per algorithmic block)"),
        "tc": ("hgsfa::layer_tc_kernel (tcgen05 kind::tf32)", "tensor", bf16 / 6.0,
               bf16_txt + " / 6 (TF32 = bf16 / 2; 3xTF32 split = 3 MMAs per algorithmic block)"),
        "ffma": ("hgsfa::layer_kernel (packed FP32 FMA)", "fp32", fp32_peak, fp32_src)}
    kernels = []
    for name, e in eng.items():
        kname, bound, peak, src = ceilings[name]
        a = e["achieved"]
        kernels.append({"kernel": kname, "ops": e["ops"], "ms_per_step": e["ms"], "bound": bound, "achieved": a, "peak": peak,
                        "unit": "TFLOP/s", "frac": a / peak if a else None, "peak_source": src,
                        "algorithmic_flops_per_window": e["alg"] / n, "executed_pipe_flops_per_step": e["exe"]})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    dom = kernels[0] if kernels else None
    out = {}
    if dom:
        out = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": "TFLOP/s", "frac": dom["frac"],
               "kernel": "%s: %d of %d layer ops, %.1f of %.1f ms per step" % (dom["kernel"], dom["ops"], len(op_stats),
                                                                              dom["ms_per_step"], kernel_ms_last),
               "peak_source": dom["peak_source"], "kernels": kernels}
    whole = fl["algorithmic"] / (kernel_ms_last * 1e-3) / 1e12 if kernel_ms_last > 0 else None
    out.update({
        "traffic": (traffic * n) if traffic else None, "traffic_unit": "DRAM bytes per step, all layer launches (ncu)",
        "traffic_source": tsrc, "kernel_share_of_step_ncu": lshare,
        "algorithmic_bytes_per_step": fl["min_bytes"],
        "algorithmic_flops_per_window": fl["algorithmic"] / n,
        "whole_step_algorithmic_tflops": whole, "kernel_ms_per_step": kernel_ms_last,
        "per_op_ms": [round(st["ms"] / steps, 3) for st in op_stats],
        "per_op_engine": [st["engine"] for st in op_stats],
        "hbm_frac": (fl["min_bytes"] / (kernel_ms_last * 1e-3) / 1e9 / peaks["hbm_gbs"]) if kernel_ms_last > 0 else None,
        "hbm_peak_source": peaks["_source"]})
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        # started before the warm-up (nvidia-smi needs up to a second to deliver its first sample on an 8-GPU box);
        # stop(t0, t1) keeps the samples whose timestamp falls inside the timed region
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def has_output(self):
        try:
            return self.proc is None or os.path.getsize(self.path) > 0
        except OSError:
            return True

    def stop(self, t0=None, t1=None):
        """t0, t1: wall-clock (time.time()) bounds of the timed region; samples outside are dropped unless none
        fall inside, in which case the samples taken under load since the warm-up are used and the window says so."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            with open(self.path) as f:
                for line in f:
                    c = [v.strip() for v in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        rows.append((ts, float(c[1]), float(c[2]),
                                     [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                               "sw_power_cap"), c[5:9]) if v.lower().startswith("active")]))
                    except ValueError:
                        continue
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if t0 is None or (t0 - 0.05 <= r[0] <= t1 + 0.05)]
        window = "timed region"
        if not inside and rows:
            inside = [r for r in rows if r[1] > 0.9 * max(x[1] for x in rows)] or rows     # under load since the warm-up
            window = "warm-up + timed region (no nvidia-smi sample fell inside the timed region)"
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted({n for r in inside for n in r[3]}), samples=len(inside), window=window)
        return out


# ------------------------------------------------------------------------------------------------------------------
# detection leg: BASELINE configs[2] (default) / configs[4] (--detect-config 4)
# ------------------------------------------------------------------------------------------------------------------
DETECT_CONFIGS = {
    2: dict(name="configs[2]", hw=(1080, 1920), smallest_face=0.05, prescale=True, images=64),
    4: dict(name="configs[4]", hw=(2160, 3840), smallest_face=0.02, prescale=False, images=21),
}


class _StageTimes(object):
    """The add_task_ellapsed surface of the reference's benchmarking.Benchmark: device seconds per label."""

    def __init__(self):
        self.tasks = {}

    def add_task_ellapsed(self, task_label, ellapsed_time, reference=None):
        t, k = self.tasks.get(task_label, (0.0, 0))
        self.tasks[task_label] = (t + ellapsed_time, k + 1)


def _detect_scenes(cm, cfg, n_bases=8):
    """Deterministic synthetic scenes (smooth noise + face-like blobs); a rank's images are shifted copies of these."""
    H, W = cfg["hw"]
    bases = []
    for b in range(n_bases):
        rng = np.random.default_rng(9000 + b)
        s = min(H, W)
        faces = [(rng.uniform(0.1 * W, 0.9 * W), rng.uniform(0.15 * H, 0.85 * H), rng.uniform(0.07 * s, 0.3 * s), rng.uniform(-10, 10))
                 for _ in range(4)]
        bases.append(cm.render_scene(H, W, faces, 9100 + b))
    return bases


def run_detect(args, rank, world, local_rank, dev, barrier, dist):
    """images/s of the full detection path on synthetic images, sharded by image over the ranks (weak scaling: every rank
    owns `images` images).  Timed region per step: pinned host images -> device, NEAREST prescale (FaceDetectUpdated.py:551),
    window pyramid, 17 face stages, eye stage, host purge, and the gather of the per-image detection lists."""
    import torch
    from threadpoolctl import threadpool_limits
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cascade_models as cm                       # synthetic model set (flows + heads); cached under build/flows
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, shard
    from pyfaceanalysis_b200.cascade import FaceDetector
    cfg = DETECT_CONFIGS[args.detect_config]
    m = cm.cached_models(spec=FLOW_SPEC)
    all_flows = []

    def model_set():
        """One set of device objects (plans with their workspaces, heads): a lane owns its own, nothing is shared."""
        flows, heads = {}, {}
        nets = [None if f is None else flows.setdefault(id(f), GpuFlow(f, device=local_rank)) for f in m["networks"]]
        clfs = [None if c is None else heads.setdefault(id(c), GpuGaussianClassifier(c, device=local_rank)) for c in m["classifiers"]]
        all_flows.extend(flows.values())
        return nets, clfs
    nets, clfs = model_set()
    n_img = args.detect_images or cfg["images"]
    total = n_img * world
    mine = shard.image_shard(total, rank, world)                     # round robin over the global image list
    base = _detect_scenes(cm, cfg, min(8, n_img))
    host = [torch.from_numpy(np.ascontiguousarray(np.roll(base[k % len(base)], (k * 37) % cfg["hw"][1], axis=1))).pin_memory()
            for k in mine]                                            # image k of the global list: base k % 8 shifted by 37 k columns
    # The synthetic heads are not trained to the reference's operating point: calibrate the Disc cut-offs on one image
    # (the same on every rank) so that the funnel has the shape a cascade is built for: few windows survive Disc1.
    keep = {"Disc1": 0.04, "Disc3": 0.4, "Disc5": 0.5, "Disc7": 0.6, "Disc9": 0.5}
    cut = [1e30] * 10
    cal = FaceDetector(m["header"], m["network_types"], nets, clfs, cut_offs_face=cut, header_eye=None, device=local_rank)
    cal_img = cal.prescale([torch.from_numpy(base[0]).to(dev)])[0] if cfg["prescale"] else torch.from_numpy(base[0]).to(dev)
    for name, frac in keep.items():
        cal.cut_offs = list(cut)
        _, tr0 = cal.detect([cal_img], smallest_face=cfg["smallest_face"], return_trace=True)
        sc = tr0["disc_scores"].get(name)
        cut[int(name[-1])] = float(np.nanquantile(sc, frac)) if sc is not None and len(sc) else 0.0
    # Throughput mode of the detector: `lanes` independent FaceDetector instances, each driven by its own host thread on
    # its own CUDA stream, take alternate batches, so that the host part of one batch (launch loop, eye bookkeeping, purge)
    # runs under the kernels of the other lane's batch; the pinned-host -> device copy of a lane's next batch runs on a copy
    # stream, and the gather of the detection lists (host objects, gloo) on a helper thread, in batch order.  Every batch is
    # still uploaded, processed and gathered exactly once per step, all inside the timed region.
    import threading
    from concurrent.futures import ThreadPoolExecutor
    n_lanes = max(1, args.detect_lanes)
    lanes = []
    for j in range(n_lanes):
        ln_nets, ln_clfs = (nets, clfs) if j == 0 else model_set()
        lanes.append(dict(det=FaceDetector(m["header"], m["network_types"], ln_nets, ln_clfs, cut_offs_face=cut,
                                           header_eye=m["header_eye"], device=local_rank),
                          stream=torch.cuda.Stream(dev), up=torch.cuda.Stream(dev)))
    det = lanes[0]["det"]
    gather_group = dist.new_group(backend="gloo") if world > 1 else None
    pool = ThreadPoolExecutor(1)

    def upload(lane):
        with torch.cuda.stream(lane["up"]):
            imgs = [h.to(dev, non_blocking=True) for h in host]
            ev = torch.cuda.Event()
            ev.record(lane["up"])
        return imgs, ev

    def lane_loop(j, k, n_steps, out, done, bench, errors):
        """Lane j of k active lanes: batches j, j + k, ... on this thread and this lane's stream."""
        lane = lanes[j]
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(lane["stream"]):
                mine_steps = list(range(j, n_steps, k))
                nxt = upload(lane) if mine_steps else None
                for pos, i in enumerate(mine_steps):
                    imgs, ev = nxt
                    lane["stream"].wait_event(ev)
                    for t in imgs:
                        t.record_stream(lane["stream"])
                    nxt = upload(lane) if pos + 1 < len(mine_steps) else None
                    if cfg["prescale"]:
                        imgs = lane["det"].prescale(imgs)
                    out[i] = lane["det"].detect(imgs, smallest_face=cfg["smallest_face"], return_trace=True, benchmark=bench)
                    done[i].set()
        except BaseException as e:          # noqa: BLE001 -- re-raised by run()
            errors.append(e)
            for d in done:
                d.set()

    def run(n_steps, bench=None, use_lanes=None):
        k = n_lanes if use_lanes is None else use_lanes
        out, done, errors = [None] * n_steps, [threading.Event() for _ in range(n_steps)], []
        threads = [threading.Thread(target=lane_loop, args=(j, k, n_steps, out, done, bench, errors)) for j in range(k)] if k > 1 else []
        for t in threads:
            t.start()
        if k == 1:
            lane_loop(0, 1, n_steps, out, done, bench, errors)
        pending, allr = None, None
        for i in range(n_steps):                                       # gathers in batch order: same order on every rank
            done[i].wait()
            if errors:
                break
            if pending is not None:
                allr = pending.result()
            pending = pool.submit(shard.gather_detections, out[i][0], mine, total, gather_group)
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        allr = pending.result()
        return allr, out[-1][1]

    steps = max(n_lanes, min(max(args.steps, 6), 12) // n_lanes * n_lanes)      # 6..12 batches, a multiple of the lanes
    allr, tr = run(max(n_lanes, min(args.warmup, 2) * n_lanes))
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    allr, tr = run(steps)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item()) / steps
    stage_bench = _StageTimes()
    run(1, stage_bench, use_lanes=1)                                  # one more, untimed, alone: device times per label
    pool.shutdown()
    res = None
    if rank == 0:
        peaks = _peaks()
        crops = [int(tr["stage_counts"][k]) for k in range(len(det.types)) if det.networks[k] is not None and
                 not (k > 0 and det.types[k - 1][:-1] == "Disc")]
        crop_s = stage_bench.tasks.get("Extraction of subimages patches", (0.0, 0))[0]
        crop_bytes = float(sum(crops)) * 4096 * 2
        res = {"metric": "detect images/sec", "value": total / dt, "unit": "images/s", "ms_per_batch": dt * 1e3, "n_gpus": world,
               "scaling": "weak", "images_per_gpu": n_img, "images_total": total, "steps": steps,
               "config": {"workload": "%s: %d synthetic %dx%d 'L' images per GPU, smallest_face %.2f, %s, synthetic %s cascade "
                                      "(17 face stages + eye stage + purge)" % (cfg["name"], n_img, cfg["hw"][1], cfg["hw"][0],
                                                                               cfg["smallest_face"], "NEAREST prescale to <= 1000 px on the device"
                                                                               if cfg["prescale"] else "no prescale (--image_prescaling=0)", FLOW_SPEC),
                          "sharding": "images round robin over ranks (shard.image_shard), detection lists gathered with "
                                      "all_gather_object on a gloo group (shard.gather_detections)",
                          "lanes": n_lanes,
                          "overlap": "%d detector instances on their own streams and host threads take alternate batches; upload of a "
                                     "lane's next batch on a copy stream, gather of the detection lists on a helper thread; "
                                     "stage_ms is one batch alone on one lane" % n_lanes},
               "windows_per_gpu": int(tr["n_windows"]), "stage_counts": [int(c) for c in tr["stage_counts"]],
               "detections_total": int(sum(len(o) for o in allr)), "host_syncs_per_batch": int(tr["host_syncs"]),
               "calibrated_cut_offs": cut,
               "e2e": {"value": total / dt, "unit": "images/s", "h2d_bytes_per_step": int(sum(h.numel() for h in host)),
                       "d2h_bytes_per_step": int(sum(len(o) for o in allr) * 80 // max(world, 1)),
                       "api": "FaceDetector.prescale + FaceDetector.detect on pinned host images"},
               "stage_ms": {k: round(v[0] * 1e3, 3) for k, v in stage_bench.tasks.items()},
               "crop_roofline": {"bound": "hbm", "kernel": "hgsfa::crop_rows_u8_kernel (row-major uint8 patches)",
                                 "windows_cropped": int(sum(crops)), "achieved": crop_bytes / crop_s / 1e9 if crop_s > 0 else None,
                                 "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": crop_bytes / crop_s / 1e9 / peaks["hbm_gbs"] if crop_s > 0 else None,
                                 "bytes_per_window": "4096 gathered + 4096 written (SURVEY.md 8d)", "peak_source": peaks["_source"]}}
        if not args.no_cpu_baseline and world == 1:
            from oracle import cascade as ocascade
            threads = min(12, os.cpu_count() or 1)
            img0 = cal_img.cpu().numpy()
            with threadpool_limits(limits=threads):
                c0 = time.perf_counter()
                ocascade.detect_image(img0, m["header"], m["network_types"], m["networks"], m["classifiers"], cfg["smallest_face"],
                                      m["num_face_stages"], cut_offs_face=cut, eye_header=m["header_eye"])
                cdt = time.perf_counter() - c0
            res["cpu_baseline"] = {"value": 1.0 / cdt, "unit": "images/s", "cores": threads, "nproc": os.cpu_count(), "kind": "port",
                                   "sample": "1 image (%d windows), float64 numpy oracle of the reference loop, %.1f s"
                                             % (int(tr["n_windows"]) // n_img, cdt)}
    for g in all_flows:
        g.close()
    return res



def cpu_oracle_windows_per_s(flow, n_sample, threads, reps=1):
    """The float64 numpy restatement (oracle/) on host cores: the 'port' CPU baseline."""
    from oracle import nodes as onodes
    from threadpoolctl import threadpool_limits
    rng = np.random.default_rng(SEED)
    x = rng.integers(0, 256, (n_sample, 4096), dtype=np.uint8).astype(np.float64)
    with threadpool_limits(limits=threads):
        onodes.flow_execute(flow, x[:256])      # warm caches / BLAS threads
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            onodes.flow_execute(flow, x)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_sample / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm for this path (oracle port) on the host cores."""
    if rank != 0:
        return
    from pyfaceanalysis_b200 import synthetic
    flow = synthetic.cached_flow(FLOW_SPEC, seed=0)
    threads = min(12, os.cpu_count() or 1)
    n_sample = 4096
    from oracle import nodes as onodes
    from threadpoolctl import threadpool_limits
    rng = np.random.default_rng(SEED)
    x = rng.integers(0, 256, (n_sample, 4096), dtype=np.uint8).astype(np.float64)
    with threadpool_limits(limits=threads):
        for _ in range(max(1, args.warmup)):
            onodes.flow_execute(flow, x[:512])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            onodes.flow_execute(flow, x)
        dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = "%d windows x 4096 px per step (uniform uint8, seed %d), float64 numpy oracle" % (n_sample, SEED)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "U11L_64 flow forward, bounded CPU sample of configs[1]", "windows_per_step": n_sample,
                   "window_dim": 4096},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "nproc": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--windows", type=int, default=N_WINDOWS, help="windows per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--front-chunk", type=int, default=0, help="windows per front-segment chunk (0 = library default)")
    ap.add_argument("--back-chunk", type=int, default=0)
    ap.add_argument("--no-detect", action="store_true", help="skip the detection leg (BASELINE configs[2])")
    ap.add_argument("--detect-config", type=int, default=2, choices=sorted(DETECT_CONFIGS), help="2: 64 x 1920x1080 prescaled; "
                    "4: 21 x 3840x2160 without prescale (~1e6 windows per GPU)")
    ap.add_argument("--detect-images", type=int, default=0, help="images per GPU (0 = the config's)")
    ap.add_argument("--detect-lanes", type=int, default=2, help="detector instances (streams + host threads) per GPU taking alternate batches")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pyfaceanalysis_b200 import GpuFlow, _lib, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    flow = synthetic.cached_flow(FLOW_SPEC, seed=0)
    g = GpuFlow(flow, device=local_rank)
    if args.front_chunk or args.back_chunk:
        g.set_chunks(front=args.front_chunk, back=args.back_chunk)
    n = args.windows
    F = g.output_dim

    # ---- synthetic input, resident in HBM, row-major (N, 4096) uint8 ----
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + rank)
    x_dev = torch.randint(0, 256, (n, g.input_dim), dtype=torch.uint8, device=dev, generator=gen)
    y_dev = torch.empty((n, F), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        g.execute_torch(x_dev, out=y_dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    # keep the GPU loaded (extra untimed steps) until nvidia-smi has delivered its first sample, 3 s at most
    t_wait = time.time()
    while not sampler.has_output() and time.time() - t_wait < 3.0:
        step()
        torch.cuda.synchronize(dev)
    g.profile(True)          # CUDA events around every layer launch of the timed steps (per-op device time)
    l0 = g.stats()["launches"]
    barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step_kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    clocks = sampler.stop(wall0, time.time())
    ms_total = e0.elapsed_time(e1)
    launches = g.stats()["launches"] - l0
    kernel_ms_last = g.stats()["last_ms"]      # device time of the last step's launches (plan events)
    op_stats = g.op_stats()                    # per-op totals over the timed steps
    g.profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers ----
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((n, g.input_dim), dtype=torch.uint8).pin_memory()
        x_host.copy_(x_dev)
        y_host = torch.empty((n, F), dtype=torch.float64).pin_memory()
        xh, yh = x_host.numpy(), y_host.numpy()
        g.execute(xh, out=yh)
        barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            g.execute(xh, out=yh)    # returns after the D2H copy of the features has completed
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(n * g.input_dim),
               "d2h_bytes_per_step": int(n * F * 8), "steps": e2e_steps,
               "api": "GpuFlow.execute(x_host_pinned_u8, out=y_host_pinned_f64)"}
        # keep the parity honest inside the bench as well: device path == host path
        chk = y_dev[:4096].double().cpu().numpy()
        if not np.allclose(chk, yh[:4096], rtol=1e-5, atol=1e-3):
            raise SystemExit("bench: device-resident and host-path results differ")
        del x_host, y_host

    detect = None
    if not args.no_detect:
        del x_dev, y_dev
        torch.cuda.empty_cache()
        detect = run_detect(args, rank, world, local_rank, dev, barrier, dist if world > 1 else None)

    if rank == 0:
        fl = g.flops(n)
        fp32_peak, fp32_src = _fp32_peak()
        peaks = _peaks()
        tpw, tsrc, lshare = _traffic_per_window()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "2-piece f16 split (fused front and bounded layer ops) / 3xtf32 (unbounded inputs, HGSFA_TC_F16=0), f32 accumulation; f32 FFMA with HGSFA_ENGINE=ffma", "data": "synthetic",
            "config": {"workload": "configs[1]: U11L_64 (FaceCentering2-shaped synthetic HiGSFA flow) forward, "
                                   "%d windows x 4096 uint8 per GPU per step" % n,
                       "windows_per_gpu": n, "window_dim": g.input_dim, "features": F, "flow": FLOW_SPEC,
                       "input_layout": "row-major (N,4096) uint8 resident in HBM",
                       "l2": "inputs (%.1f GB/step) far larger than the 126 MB L2; no flush needed" % (n * 4096 / 1e9),
                       "parallelism": "windows sharded across GPUs, no collective"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "samples": clocks["samples"], "window": clocks.get("window")},
            "roofline": _roofline(op_stats, args.steps, n, fl, kernel_ms_last, peaks, fp32_peak, fp32_src, tpw, tsrc, lshare),
        }
        if detect is not None:
            out["detect"] = detect
        if not args.no_cpu_baseline:
            threads = min(12, os.cpu_count() or 1)
            n_sample = 32768
            v, secs = cpu_oracle_windows_per_s(flow, n_sample, threads)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "nproc": os.cpu_count(), "kind": "port",
                                   "sample": "%d windows x 4096 px (same distribution), float64 numpy oracle, %.1f s"
                                             % (n_sample, secs)}
        print(json.dumps(out))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
