"""Builder tool: libhgsfa variants with extra -D flags for one translation unit (kernel experiments).

    python tools/build_variant.py NAME front.cu -DHGSFA_FR_SLEEP_MMA=0 ...   ->  build/variants/libhgsfa_NAME.so

The other objects are taken from the last in-tree build (build/obj); select a variant at run time with HGSFA_LIB=path."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfaceanalysis_b200 import build as B  # noqa: E402

name, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "%s_%s.o" % (unit.replace(".cu", ""), name))
subprocess.run([B.nvcc_path()] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, unit), "-o", obj], check=True)
objs = [obj if s == unit else os.path.join(ROOT, "build", "obj", s.replace(".cu", ".o")) for s in B.SOURCES]
lib = os.path.join(out_dir, "libhgsfa_%s.so" % name)
subprocess.run([B.nvcc_path(), "-shared", "-o", lib] + objs + ["-lcudart"], check=True)
print(lib)
