#!/usr/bin/env python
"""Process-level timing of HYMET's whole Mash stage (run_hymet_cami.sh:85-97) on C2-shaped data split into
three sketch files: (a) bin/hymet-mash-stage, one pass; (b) three `mash.sh`-equivalent runs with bin/mash
(mash screen + the post-processing, one process per sketch file).  Files in the page cache; wall clock.
  python tools/stage_bench.py [--mbp 1000]
"""
import argparse, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=int, default=1000)
    ap.add_argument("--dir", default="/tmp/hs_stage_bench")
    a = ap.parse_args()
    from hymet_b200 import msh as mshfmt, stage, synth, workload
    os.makedirs(os.path.join(a.dir, "input"), exist_ok=True)
    os.makedirs(os.path.join(a.dir, "out"), exist_ok=True)
    wl = workload.make_c2(0, mbp=a.mbp, n_sketches=50_000, n_real=500, with_fasta=True, with_host_packed=False)
    n = len(wl.lengths)
    order = np.random.default_rng(1).permutation(n)          # real genomes spread over the three files
    parts = [order[:20_000], order[20_000:40_000], order[40_000:]]
    dbs = []
    H = wl.hashes.reshape(n, wl.s)
    for j, idx in enumerate(parts):
        idx = np.sort(idx)
        db = mshfmt.SketchDB(k=wl.k, s=wl.s, names=[synth.gcf_name(int(i)) for i in idx], comments=["synthetic %d" % i for i in idx],
                             lengths=wl.lengths[idx], offsets=np.arange(len(idx) + 1, dtype=np.uint64) * np.uint64(wl.s),
                             hashes=H[idx].reshape(-1))
        p = os.path.join(a.dir, "sketch%d.msh" % (j + 1))
        mshfmt.write_msh(p, db)
        dbs.append(p)
    wl.fasta.numpy().tofile(os.path.join(a.dir, "input", "sample_0.fna"))
    del wl
    names = ("screen.tab", "filtered.tab", "sorted.tab", "top_hits.tab", "selected.txt")
    outs = [[os.path.join(a.dir, "out", "%d_%s" % (j, f)) for f in names] for j in range(3)]
    res = {"fasta_bytes": os.path.getsize(os.path.join(a.dir, "input", "sample_0.fna")), "msh_bytes": [os.path.getsize(p) for p in dbs]}
    py = sys.executable
    # (a) one pass
    times = []
    for rep in range(3):
        argv = [py, os.path.join(ROOT, "bin", "hymet-mash-stage"), "--merge", "-p", "8", os.path.join(a.dir, "input"), "0.9"]
        for p, o in zip(dbs, outs):
            argv += [p] + o
        t0 = time.perf_counter()
        r = subprocess.run(argv, capture_output=True)
        times.append(time.perf_counter() - t0)
        assert r.returncode == 0, r.stderr.decode()[-500:]
    res["fused_stage_wall_s"] = times
    fused = [[open(f, "rb").read() for f in o] for o in outs]
    # (b) three processes + post-processing
    times = []
    for rep in range(3):
        t0 = time.perf_counter()
        sel = []
        for j, p in enumerate(dbs):
            r = subprocess.run([py, os.path.join(ROOT, "bin", "mash"), "screen", "-p", "8", "-v", "0.9", p,
                                os.path.join(a.dir, "input", "sample_0.fna")], capture_output=True)
            assert r.returncode == 0
            s = stage.select(r.stdout, 1, "0.9")
            sel.append(s["selected"])
            if rep == 0:
                assert r.stdout == fused[j][0] and s["sorted"] == fused[j][2] and s["top_hits"] == fused[j][3]
        merged = stage.merge_selected(sel)
        times.append(time.perf_counter() - t0)
    res["three_screens_wall_s"] = times
    res["identical_outputs"] = bool(merged == fused[0][4])
    res["selected_genomes"] = merged.count(b"\n")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
