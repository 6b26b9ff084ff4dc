"""Builder tool: per-phase cycle counts of expansion warp 0 of one CTA of layer 3 (needs a library built with
-DHGSFA_TC_TRACE=32: python tools/build_variant.py trace flow.cu -DHGSFA_TC_TRACE=32; HGSFA_LIB=build/variants/libhgsfa_trace.so)."""
import ctypes as C
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import GpuFlow, _lib, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
g = GpuFlow(synthetic.cached_flow("U11L_64", seed=0), device=0)
x = torch.randint(0, 256, (n, 4096), dtype=torch.uint8, device="cuda")
out = torch.empty((n, 60), dtype=torch.float32, device="cuda")
for _ in range(3):
    g.execute_torch(x, out=out)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 16384)()
lib.hgsfa_debug_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.hgsfa_debug_trace(buf, 16384) == 0
cnt = int(buf[0])
ev = [(int(v) >> 8, int(v) & 0xff) for v in list(buf)[1:cnt]]
print("events", len(ev))
names = {(0, 1): "wait XFULL (node start)", (1, 2): "bias copy + node setup", (2, 3): "wait AFREE", (3, 4): "segments of a chunk (compute + tcgen05.st)",
         (4, 5): "tcgen05.wait::st + fence", (5, 6): "syncwarp + arrive AFULL", (6, 2): "loop to next chunk", (6, 7): "node end: arrive XFREE",
         (7, 0): "loop to next node", (8, 9): "16-term iteration: loads + arithmetic", (9, 10): "16-term iteration: split + 4 x tcgen05.st",
         (3, 8): "chunk start -> first 16-term iteration", (10, 8): "between 16-term iterations", (10, 4): "last 16-term iteration -> end of chunk"}
tot = defaultdict(int)
num = defaultdict(int)
for (t0, s0), (t1, s1) in zip(ev[:-1], ev[1:]):
    tot[(s0, s1)] += t1 - t0
    num[(s0, s1)] += 1
nodes = num[(0, 1)]
span = ev[-1][0] - ev[0][0]
print("nodes traced %d, %.0f cycles per node (one tile per node)" % (nodes, span / max(nodes, 1)))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("  %-45s %6.1f %% of the span, %8.0f cycles per node, %7.0f per occurrence (%d)" % (
        names.get(k, str(k)), 100.0 * tot[k] / span, tot[k] / max(nodes, 1), tot[k] / num[k], num[k]))
