#!/usr/bin/env python
"""SURVEY.md 8f rank 3, the measurement behind the design: what does the fixed cost of one `mash screen`
process consist of, and would a cache file of the BUILT table beat rebuilding it?

For a sketch database of the C2 / C3 shape (decoy sketches: parse and build cost do not depend on the
hash values) this times, on the GPU box, files in the page cache:
  parse_s + build_s   hs_msh_open (Cap'n Proto walk) + table build on the GPU = what every process pays today
  cache_*             reading a file of the built table's size back into HBM, two ways:
                      mmap + one copy from pageable memory; pread into pinned chunks + async copies
  context_s           CUDA context creation alone (a fresh process), which neither form avoids
The resident server (hymet_b200/server.py) avoids all three; its attach time is in tools/cli_bench.py.
  python tools/f3_cache_bench.py [--sketches 50000] [--dir /tmp/hs_f3]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sketches", type=int, default=50_000)
    ap.add_argument("--s", type=int, default=1000)
    ap.add_argument("--dir", default="/tmp/hs_f3")
    a = ap.parse_args()
    import torch

    from hymet_b200 import msh as mshfmt, screen as hs, synth
    os.makedirs(a.dir, exist_ok=True)
    rng = np.random.default_rng(7)
    decoy, dlen = synth.decoy_sketches(rng, a.sketches, a.s)
    n = a.sketches
    db = mshfmt.SketchDB(k=21, s=a.s, names=[synth.gcf_name(i) for i in range(n)], comments=["synthetic %d" % i for i in range(n)],
                         lengths=dlen, offsets=np.arange(n + 1, dtype=np.uint64) * np.uint64(a.s), hashes=decoy.reshape(-1))
    dbp = os.path.join(a.dir, "db_%d.msh" % n)
    mshfmt.write_msh(dbp, db)
    del db, decoy
    out = {"sketches": n, "s": a.s, "msh_bytes": os.path.getsize(dbp)}
    # fresh process: CUDA context creation alone
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from hymet_b200 import _abi; _abi.init(0)" % ROOT], check=True)
    out["fresh_process_context_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from hymet_b200 import _abi; _abi.load()" % ROOT], check=True)
    out["fresh_process_no_context_s"] = time.perf_counter() - t0
    # what a process pays today (parse overlaps context creation in the CLI; here they are separate numbers)
    runs = []
    for _ in range(3):
        t0 = time.perf_counter()
        d = hs.Database.load_msh(dbp)
        wall = time.perf_counter() - t0
        runs.append({"wall_s": wall, "parse_s": d.info.t_parse_s, "build_s": d.info.t_build_s})
        table_bytes = int(d.info.device_bytes)
        d.close()
    out["parse_and_build"] = runs
    out["built_table_bytes"] = table_bytes
    # a cache file of that size, read back
    cache = os.path.join(a.dir, "cache_%d.bin" % n)
    chunk = np.random.default_rng(1).integers(0, 255, size=64 << 20, dtype=np.uint8)
    with open(cache, "wb") as fh:
        left = table_bytes
        while left > 0:
            fh.write(chunk[:min(left, len(chunk))].tobytes())
            left -= len(chunk)
    with open(cache, "rb") as fh:      # warm the page cache
        while fh.read(256 << 20):
            pass
    dev = torch.device("cuda", 0)
    dst = torch.empty(table_bytes, dtype=torch.uint8, device=dev)
    mm = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        src = torch.from_file(cache, shared=False, size=table_bytes, dtype=torch.uint8)
        dst.copy_(src)
        torch.cuda.synchronize()
        mm.append(time.perf_counter() - t0)
        del src
    out["cache_mmap_copy_s"] = mm
    pin = [torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    pr = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with open(cache, "rb", buffering=0) as fh:
            off, i = 0, 0
            while off < table_bytes:
                b = pin[i & 1]
                evs[i & 1].synchronize()
                nread = fh.readinto(memoryview(b.numpy())[:min(len(b), table_bytes - off)])
                dst[off:off + nread].copy_(b[:nread], non_blocking=True)
                evs[i & 1].record()
                off += nread
                i += 1
        torch.cuda.synchronize()
        pr.append(time.perf_counter() - t0)
    out["cache_pread_pinned_s"] = pr
    best_rebuild = min(r["parse_s"] + r["build_s"] for r in runs)
    out["verdict"] = {"rebuild_s": best_rebuild, "best_cache_read_s": min(mm + pr),
                      "cache_wins": min(mm + pr) < best_rebuild,
                      "note": "either way a fresh process still pays interpreter start + CUDA context creation; only a resident "
                              "process (hymet_b200/server.py) removes the fixed cost"}
    os.remove(cache)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
