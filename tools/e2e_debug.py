#!/usr/bin/env python
"""Scratch: where does the end-to-end time go (host pack / lock / feed), and fast-path A/B."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hymet_b200 import screen as hs, workload

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = workload.make_c2(0, mbp=mbp, n_sketches=50000, n_real=64, with_fasta=True, with_host_packed=True)
db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
scr = hs.Screen(db)
nthreads = os.cpu_count()
print("cpus", nthreads, "text bytes", wl.fasta.numel())
for fast in (0,):
    for rep in range(3):
        scr.reset(); scr.feed_packed_device(wl.d_seq.data_ptr(), wl.d_inv.data_ptr(), wl.n_positions); r = scr.finish()
    print("fast=%d resident: stream %.3f ms (%.1f Gbp/s) reduce %.3f" % (fast, r.stats["ms_stream"], wl.n_bases / r.stats["ms_stream"] / 1e6, r.stats["ms_reduce"]))
for thr in (nthreads, nthreads // 2, 4, 1):
    for rep in range(2):
        scr.reset()
        t0 = time.perf_counter()
        scr.feed_text_ptr(wl.fasta.data_ptr(), wl.fasta.numel(), thr)
        t1 = time.perf_counter()
        r = scr.finish()
        t2 = time.perf_counter()
    print("threads=%d feed %.1f ms finish %.1f ms -> %.1f Gbp/s  (launches %d, stream %.2f ms)" % (
        thr, 1e3 * (t1 - t0), 1e3 * (t2 - t1), wl.n_bases / (t2 - t0) / 1e9, r.stats["n_launches"], r.stats["ms_stream"]))
# host packer alone on the pinned text, per thread count
import ctypes as C, threading
L = hs._abi.load()
txt = wl.fasta.numpy()
n = txt.size
def pack_range(b, e, out):
    cap = (e - b) // 32 + 4
    seq = np.empty(cap, np.uint64); inv = np.empty(cap, np.uint32); nb = C.c_uint64()
    t0 = time.perf_counter()
    L.hs_pack_text(C.c_void_p(txt.ctypes.data + b), e - b, C.c_void_p(seq.ctypes.data), C.c_void_p(inv.ctypes.data), cap, C.byref(nb), None)
    out.append(time.perf_counter() - t0)
for thr in (1, 4, nthreads):
    outs = []; ths = []
    t0 = time.perf_counter()
    for i in range(thr):
        ths.append(threading.Thread(target=pack_range, args=(i * n // thr, (i + 1) * n // thr, outs)))
    [t.start() for t in ths]; [t.join() for t in ths]
    dt = time.perf_counter() - t0
    print("pack only, %d threads: %.1f ms wall -> %.2f GB/s aggregate (per-thread busy %.1f ms)" % (thr, 1e3 * dt, n / dt / 1e9, 1e3 * max(outs)))
if os.environ.get("HYMET_SCREEN_DEBUG_TIMING"):
    scr.reset(); scr.feed_text_ptr(wl.fasta.data_ptr(), wl.fasta.numel(), nthreads); scr.finish()
