"""Builder tool: per-op times of the U11L_64 flow on resident uint8 windows (one line; used to compare library variants)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import GpuFlow, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
reps = 3
g = GpuFlow(synthetic.cached_flow("U11L_64", seed=0), device=0)
x = torch.randint(0, 256, (n, 4096), dtype=torch.uint8, device="cuda")
out = torch.empty((n, 60), dtype=torch.float32, device="cuda")
for _ in range(2):
    g.execute_torch(x, out=out)
torch.cuda.synchronize()
g.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    g.execute_torch(x, out=out)
e1.record()
torch.cuda.synchronize()
st = g.op_stats()
print("%s: %.3f ms/step  per-op %s  checksum %.6g" % (os.environ.get("HGSFA_LIB", "default").split("/")[-1], e0.elapsed_time(e1) / reps,
                                                      " ".join("%.2f" % (o["ms"] / reps) for o in st), float(out.double().sum())))
