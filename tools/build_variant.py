#!/usr/bin/env python
"""Kernel experiments: an alternative build of the same sources with extra -D flags, loaded through
HYMET_SCREEN_LIB (hymet_b200/_abi.py).  python tools/build_variant.py NAME -DHS_ILP=2 ... -> gpurun_variants/libhs_NAME.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hymet_b200 import build as b
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "gpurun_variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, "libhs_%s.so" % name)
cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + flags + [os.path.join(b.CSRC, f) for f in b.SOURCES] + ["-o", out, "-lz"]
subprocess.run(cmd, check=True)
print(out)
