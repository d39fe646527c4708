#!/usr/bin/env python
"""Scratch: time the three FASTA ingest modes on pinned text (256 Mbp by default)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hymet_b200 import screen as hs, workload

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = workload.make_c2(0, mbp=mbp, n_sketches=20000, n_real=32, with_fasta=True, with_host_packed=False)
db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
scr = hs.Screen(db)
thr = os.cpu_count()
for mode in modes:
    scr.set_option("ingest", mode)
    for rep in range(reps):
        scr.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scr.feed_text_ptr(wl.fasta.data_ptr(), wl.fasta.numel(), thr)
        t1 = time.perf_counter()
        r = scr.finish()
        t2 = time.perf_counter()
    print("ingest=%d: feed %.2f ms finish %.2f ms -> %.1f Gbp/s (launches %d, h2d %.0f MB, stream %.2f ms)" % (
        mode, 1e3 * (t1 - t0), 1e3 * (t2 - t1), wl.n_bases / (t2 - t0) / 1e9, r.stats["n_launches"],
        r.stats["h2d_bytes"] / 1e6, r.stats["ms_stream"]), flush=True)
