#!/usr/bin/env python
"""Scratch: hybrid ingest (pinned FASTA text) for several device-queue depths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hymet_b200 import screen as hs, workload
wl = workload.make_c2(0, mbp=1000, n_sketches=50000, n_real=500, with_fasta=True, with_host_packed=False)
db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
ncpu = len(os.sched_getaffinity(0))
for slots, batch, chunk in ((4, 3, 16), (3, 2, 16), (2, 2, 16), (2, 3, 16), (3, 3, 16), (4, 2, 16), (4, 3, 8), (3, 4, 8), (4, 1, 16)):
    scr = hs.Screen(db)
    scr.set_option("ingest_slots", slots); scr.set_option("ingest_batch", batch); scr.set_option("chunk_bases", chunk << 20)
    ts = []
    for rep in range(6):
        scr.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
        scr.feed_text_ptr(wl.fasta.data_ptr(), wl.fasta.numel(), ncpu); t1 = time.perf_counter()
        r = scr.finish(False); ts.append((time.perf_counter() - t0, t1 - t0, r.stats["h2d_bytes"]))
    best = min(ts[2:])
    print("slots %d batch %d chunk %2d MB: best %.2f ms (feed %.2f) h2d %d MB -> %.0f Mbp/s ; all %s" % (
        slots, batch, chunk, 1e3 * best[0], 1e3 * best[1], best[2] >> 20, wl.n_bases / best[0] / 1e6, ["%.1f" % (1e3 * t[0]) for t in ts[2:]]), flush=True)
    scr.close()
