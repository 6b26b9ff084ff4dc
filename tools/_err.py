import os, sys, numpy as np
sys.path.insert(0, ".")
from pyfaceanalysis_b200 import synthetic, GpuFlow
from oracle import nodes as onodes
f = synthetic.cached_flow("U11L_64")
x = synthetic.synthetic_patches(2048, (64, 64), 123)
ref = onodes.flow_execute(f, x.astype(np.float64))
y = GpuFlow(f).execute(x, out_dtype=np.float32).astype(np.float64)
std = ref.std(axis=0)
print(os.environ.get("HGSFA_ENGINE", "auto"), "max|err|/std = %.3g" % (np.abs(y - ref) / std).max(), "median %.3g" % np.median(np.abs(y - ref) / std))
