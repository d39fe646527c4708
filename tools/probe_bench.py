#!/usr/bin/env python
"""K2 alone: random probes against a 50k-sketch table, and the raw random-sector reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hymet_b200 import screen as hs, synth

nsk = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
rng = np.random.default_rng(0)
h, ln = synth.decoy_sketches(rng, nsk, 1000)
db = hs.Database.from_arrays(21, 1000, 42, np.arange(nsk + 1, dtype=np.uint64) * 1000, h.reshape(-1), ln)
n = 1 << 26
hq = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda")
for rep in range(3):
    hits, reads, ms = db.probe_device(hq.data_ptr(), n)
    print("k_probe: %.3f ms  %.2e probes/s  reads/probe %.3f  sector GB/s %.0f" % (ms, n / ms * 1e3, reads / n, 32 * reads / ms / 1e6))
buf = torch.empty(int(db.info.n_buckets) * 4, dtype=torch.int64, device="cuda")
for rep in range(3):
    ms = hs.gather_bench(buf.data_ptr(), buf.numel() * 8, n)
    print("k_gather_bench (%.0f MB): %.3f ms  %.2e reads/s  sector GB/s %.0f" % (buf.numel() * 8 / 1e6, ms, n / ms * 1e3, 32 * n / ms / 1e6))
small = torch.empty(8 << 20, dtype=torch.int64, device="cuda")   # 64 MB: L2 resident
ms = hs.gather_bench(small.data_ptr(), small.numel() * 8, n)
print("k_gather_bench (64 MB, L2): %.3f ms  %.2e reads/s" % (ms, n / ms * 1e3))
