"""GPU check of the fused front (csrc/front_tc.cuh) against the per-layer path and the float64 oracle, plus timing.

    python tools/front_check.py [n_time]

Builder tool (run under gpurun); the tests in tests/test_gpu_flow.py cover the same ground."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import GpuFlow, _lib, synthetic  # noqa: E402
from oracle import nodes as onodes  # noqa: E402
import ctypes as C  # noqa: E402


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    flow = synthetic.cached_flow("U11L_64", seed=0)
    fused = GpuFlow(flow, device=0)
    print("fused_front:", fused.fused_front, fused.front_reason, flush=True)
    os.environ["HGSFA_FRONT"] = "0"
    plain = GpuFlow(flow, device=0)
    del os.environ["HGSFA_FRONT"]
    std = flow._train_output_std
    rng = np.random.default_rng(3)
    ok = True
    for n in (1, 128, 300, 1000):
        x = np.concatenate([synthetic.synthetic_patches(n // 2 + 1, (64, 64), 11), rng.integers(0, 256, (n, 4096), dtype=np.uint8)])[:n]
        xt = torch.as_tensor(x, device="cuda")
        y_f = fused.execute_torch(xt).cpu().numpy().astype(np.float64)
        y_p = plain.execute_torch(xt).cpu().numpy().astype(np.float64)
        ref = onodes.flow_execute(flow, x.astype(np.float64)) if n <= 300 else None
        # tiled input path
        n_pad = (n + 127) // 128 * 128
        tiled = torch.zeros(n_pad * 4096, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().hgsfa_tile_windows_device(C.c_void_p(xt.data_ptr()), _lib.U8, n, 4096, 4096, C.c_void_p(tiled.data_ptr()), _lib.U8, None))
        y_t = fused.execute_torch(tiled, layout=_lib.TILED, n=n).cpu().numpy().astype(np.float64)
        sd = np.maximum(std, y_p.std(axis=0)) if n > 1 else std      # noise rows: the batch's own spread (tests/test_gpu_flow.py)
        e_fp = np.abs(y_f - y_p).max(axis=0) / sd
        e_tp = np.abs(y_t - y_p).max(axis=0) / sd
        line = "n=%5d  fused-vs-layer %.2e  tiled-vs-layer %.2e" % (n, e_fp.max(), e_tp.max())
        if ref is not None:
            line += "  fused-vs-oracle %.2e  layer-vs-oracle %.2e" % ((np.abs(y_f - ref).max(axis=0) / sd).max(), (np.abs(y_p - ref).max(axis=0) / sd).max())
        print(line, flush=True)
        ok &= bool(e_fp.max() < 1e-3 and e_tp.max() < 1e-3)
    # host entry point (row-major through the staging buffers)
    x = rng.integers(0, 256, (70000, 4096), dtype=np.uint8)
    y_f = fused.execute(x, out_dtype=np.float32)
    y_p = plain.execute(x, out_dtype=np.float32)
    e = (np.abs(y_f.astype(np.float64) - y_p) .max(axis=0) / np.maximum(std, y_p.std(axis=0))).max()
    print("host execute 70000: fused-vs-layer %.2e" % e, flush=True)
    ok &= bool(e < 1e-3)
    # timing
    xt = torch.randint(0, 256, (n_time, 4096), dtype=torch.uint8, device="cuda")
    out = torch.empty((n_time, 60), dtype=torch.float32, device="cuda")
    for name, g in (("fused", fused), ("layer", plain)):
        for _ in range(2):
            g.execute_torch(xt, out=out)
        torch.cuda.synchronize()
        g.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            g.execute_torch(xt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        st = g.op_stats()
        g.profile(False)
        print("%s: %.3f ms per %d windows = %.2f M windows/s; per-op ms %s" % (
            name, ms, n_time, n_time / ms / 1e3, " ".join("%.2f" % (o["ms"] / reps) for o in st)), flush=True)
    print("FRONT CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
