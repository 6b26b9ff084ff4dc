"""Builder tool: a few executes of the U11L_64 flow on resident uint8 windows (target of ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import GpuFlow, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = GpuFlow(synthetic.cached_flow("U11L_64", seed=0), device=0)
x = torch.randint(0, 256, (n, 4096), dtype=torch.uint8, device="cuda")
out = torch.empty((n, 60), dtype=torch.float32, device="cuda")
for _ in range(reps):
    g.execute_torch(x, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
