#!/usr/bin/env python
"""e2e from a page-cache FASTA file for several reader counts / block sizes (scratch measurement)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hymet_b200 import screen as hs, workload

wl = workload.make_c2(0, mbp=1000, n_sketches=50000, n_real=500, with_fasta=True, with_host_packed=False)
db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
path = os.path.join(tempfile.mkdtemp(), "c.fna")
wl.fasta.numpy().tofile(path)
ncpu = len(os.sched_getaffinity(0))
for readers, block in ((4, 16), (8, 16), (12, 16), (16, 16), (8, 8), (16, 8), (8, 32)):
    scr = hs.Screen(db)
    scr.set_option("file_readers", readers)
    scr.set_option("file_block_bytes", block << 20)
    ts = []
    for rep in range(4):
        scr.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        scr.feed_fasta(path, ncpu)
        r = scr.finish(False)
        ts.append(time.perf_counter() - t0)
    print("readers %2d block %2d MB: first %.1f ms, then %s ms -> %.0f Mbp/s" % (
        readers, block, 1e3 * ts[0], ["%.1f" % (1e3 * t) for t in ts[1:]], wl.n_bases / min(ts[1:]) / 1e6), flush=True)
    scr.close()
