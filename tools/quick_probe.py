#!/usr/bin/env python
"""Scratch timing of the streaming kernel on random sequence (not the bench contract)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from hymet_b200 import screen as hs, synth

def main():
    mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsk = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    rng = np.random.default_rng(0)
    t0 = time.time()
    h, ln = synth.decoy_sketches(rng, nsk, 1000)
    offsets = (np.arange(nsk + 1, dtype=np.uint64) * 1000)
    db = hs.Database.from_arrays(21, 1000, 42, offsets, h.reshape(-1), ln)
    print("db: %d sketches, D=%d, buckets=%d, %.1f MB, build %.3fs (gen %.1fs)" % (
        db.n_refs, db.n_distinct, db.info.n_buckets, db.info.device_bytes / 1e6, db.info.t_build_s, time.time() - t0))
    n = mbp * 1_000_000
    words = hs.packed_words(n)
    seq = rng.integers(0, 2 ** 64, size=words, dtype=np.uint64)
    inv = np.zeros(words, np.uint32)
    for filt in (True, False):
        scr = hs.Screen(db, probe_filter=filt)
        for rep in range(3):
            scr.reset()
            scr.feed_packed(seq, inv, n)
            res = scr.finish(False)
            st = res.stats
            print("filter=%d rep=%d: stream %.3f ms -> %.1f Gbp/s | reduce %.3f ms | K=%d probes=%d reads=%d hits=%d mix=%d passes=%d set=%d launches=%d" % (
                filt, rep, st["ms_stream"], n / st["ms_stream"] / 1e6, st["ms_reduce"], st["n_valid_kmers"], st["n_probes"],
                st["n_bucket_reads"], st["n_hits"], st["n_mix_inserts"], st["n_mix_passes"], st["set_size"], st["n_launches"]))
        scr.close()

main()
