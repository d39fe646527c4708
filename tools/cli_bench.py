#!/usr/bin/env python
"""Process-level timing of the drop-in: `mash screen -p N -v 0.9 DB.msh contigs.fna` as HYMET calls it
(scripts/mash.sh:14), FASTA and .msh as files in the page cache, wall clock around the subprocess;
the oracle CLI on the same files beside it.  Prints one JSON object.
  python tools/cli_bench.py [--mbp 1000] [--sketches 50000] [--oracle-mbp 100]
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
    ap.add_argument("--mbp", type=int, default=1000)
    ap.add_argument("--sketches", type=int, default=50_000)
    ap.add_argument("--real", type=int, default=500)
    ap.add_argument("--oracle-mbp", type=int, default=100)
    ap.add_argument("--dir", default="/tmp/hs_cli_bench")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from hymet_b200 import msh as mshfmt, synth, workload
    os.makedirs(a.dir, exist_ok=True)
    wl = workload.make_c2(0, mbp=a.mbp, n_sketches=a.sketches, n_real=a.real, with_fasta=True, with_host_packed=False)
    n = len(wl.lengths)
    db = mshfmt.SketchDB(k=wl.k, s=wl.s, names=[synth.gcf_name(i) for i in range(n)], comments=["synthetic %d" % i for i in range(n)],
                         lengths=wl.lengths, offsets=wl.offsets, hashes=wl.hashes)
    dbp, fap, smp = [os.path.join(a.dir, f) for f in ("c2.msh", "contigs.fna", "sample.fna")]
    mshfmt.write_msh(dbp, db)
    text = wl.fasta.numpy()
    text.tofile(fap)
    # CPU sample: a record-aligned prefix
    cut = int(len(text) * a.oracle_mbp / a.mbp)
    cut = cut + int(np.argmax(text[cut:cut + (1 << 24)] == ord(">"))) if cut < len(text) else len(text)
    text[:cut].tofile(smp)
    sample_bases = int(wl.n_bases * cut / len(text))
    del wl
    out = {"query_bases": None, "fasta_bytes": os.path.getsize(fap), "msh_bytes": os.path.getsize(dbp), "sketches": n}
    env = dict(os.environ, HYMET_SCREEN_TIMING="1")
    threads = len(os.sched_getaffinity(0))
    runs = []
    for r in range(a.reps):
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", str(threads), "-v", "0.9", dbp, fap],
                           capture_output=True, env=env)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr.decode()
        runs.append({"wall_s": dt, "phases": [l for l in p.stderr.decode().splitlines() if l.startswith("[timing]")],
                     "tsv_lines": p.stdout.count(b"\n")})
        gpu_tsv = p.stdout
    out["gpu_cli"] = runs
    # the same command line against the resident table server (hymet_b200/server.py): first call builds
    # the table inside the daemon, the following ones attach to it
    sock = os.path.join(a.dir, "run", "gpu0.sock")
    senv = dict(os.environ, HYMET_SCREEN_SERVER="1", HYMET_SCREEN_SOCKET=sock)
    t0 = time.perf_counter()
    subprocess.run([sys.executable, os.path.join(ROOT, "bin", "hymet-screen-server"), "start"], check=True, capture_output=True, env=senv)
    out["server_start_s"] = time.perf_counter() - t0
    served = []
    try:
        for r in range(a.reps + 1):
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", str(threads), "-v", "0.9", dbp, fap],
                               capture_output=True, env=dict(senv, HYMET_SCREEN_TIMING="1"))
            dt = time.perf_counter() - t0
            assert p.returncode == 0, p.stderr.decode()
            served.append({"wall_s": dt, "what": "first call: table built inside the daemon" if r == 0 else "warm: table resident",
                           "phases": [l for l in p.stderr.decode().splitlines() if l.startswith("[timing]")],
                           "tsv_identical_to_in_process": p.stdout == gpu_tsv})
        t0 = time.perf_counter()
        subprocess.run([sys.executable, "-c", "pass"], check=True)
        out["python_startup_s"] = time.perf_counter() - t0
    finally:
        subprocess.run([sys.executable, os.path.join(ROOT, "bin", "hymet-screen-server"), "stop"], capture_output=True, env=senv)
    out["gpu_cli_via_server"] = served
    t0 = time.perf_counter()
    subprocess.run([sys.executable, "-c", "import numpy"], check=True)
    out["python_numpy_startup_s"] = time.perf_counter() - t0
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", str(threads), "-v", "0.9", dbp, fap],
                       capture_output=True, env=dict(env, HYMET_SCREEN_DEBUG_TIMING="1"))
    out["debug_feed_lines"] = [l for l in p.stderr.decode().splitlines() if l.startswith("[hs] block")][:12]
    from tests import _oracle as orc
    orc.build()
    t0 = time.perf_counter()
    p = subprocess.run([orc.BIN, "screen", "-p", str(threads), "-v", "0.9", dbp, smp], capture_output=True)
    out["oracle_cli"] = {"wall_s": time.perf_counter() - t0, "sample_bases": sample_bases, "threads": threads,
                         "stderr_tail": p.stderr.decode().splitlines()[-3:]}
    p2 = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", str(threads), "-v", "0.9", dbp, smp],
                        capture_output=True)
    out["sample_tsv_identical"] = p2.stdout == p.stdout and p.stdout.count(b"\n") > 0
    out["threads"] = threads
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
