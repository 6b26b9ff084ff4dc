"""Builder tool: the detector's stage-0 crop alone (64 images of 1000 x 562, smallest_face 0.05 -> 476 928 windows,
row-major uint8 patches), timed with CUDA events; target of ncu captures of crop_rows_u8_kernel."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import _lib, grid  # noqa: E402

HEADER = (40, 20, 22.5, 0.694, 0.981, 64, 64, 128, 128)
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
lib = _lib.load()
p = grid.window_pyramid(1000, 562, HEADER, 0.05, 1.1, 1.1)
coords = torch.as_tensor(np.tile(p["coords"], (n_img, 1)), device=dev)
n = coords.shape[0]
imgs = [torch.randint(0, 256, (562, 1000), dtype=torch.uint8, device=dev) for _ in range(n_img)]
ptrs = torch.tensor([t.data_ptr() for t in imgs], dtype=torch.int64, device=dev)
hw = torch.tensor([[562, 1000]] * n_img, dtype=torch.int32, device=dev)
idx = torch.repeat_interleave(torch.arange(n_img, dtype=torch.int32, device=dev), len(p["coords"]))
out = torch.empty((n, 4096), dtype=torch.uint8, device=dev)


def run():
    _lib.check(lib.hgsfa_crop_extent_batch_device(C.c_void_p(ptrs.data_ptr()), C.c_void_p(hw.data_ptr()), C.c_void_p(idx.data_ptr()),
                                                  C.c_void_p(coords.data_ptr()), None, n, 64, 64, _lib.NEAREST,
                                                  C.c_void_p(out.data_ptr()), _lib.U8, _lib.ROWMAJOR, None))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"windows": n, "ms": ms, "GB/s (4096 gathered + 4096 written per window)": n * 8192 / ms / 1e6,
                  "frac_of_6538": n * 8192 / ms / 1e6 / 6538.3}))
