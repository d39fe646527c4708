#!/usr/bin/env python
"""Host wall time of the three calls of one device-resident step (reset / feed_packed_device / finish)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hymet_b200 import screen as hs, workload

wl = workload.make_c2(0, mbp=1000, n_sketches=50000, n_real=500, with_fasta=False, with_host_packed=False)
db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
scr = hs.Screen(db, stream_ptr=stream.cuda_stream)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    scr.reset(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t1s = time.perf_counter()
    scr.feed_packed_device(wl.d_seq.data_ptr(), wl.d_inv.data_ptr(), wl.n_positions); t2 = time.perf_counter()
    torch.cuda.synchronize(); t2s = time.perf_counter()
    scr.flush(); t3 = time.perf_counter()
    r = scr.finish(False); t4 = time.perf_counter()
    print("reset call %.3f (+sync %.3f) | feed call %.3f (+sync %.3f) | flush %.3f | finish %.3f ms | ms_reduce %.3f" % (
        1e3*(t1-t0), 1e3*(t1s-t1), 1e3*(t2-t1s), 1e3*(t2s-t2), 1e3*(t3-t2s), 1e3*(t4-t3), r.stats["ms_reduce"]))
