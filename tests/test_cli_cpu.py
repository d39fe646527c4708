"""`mash` drop-in CLI: option parsing, usage/exit codes (S21) and the TSV formatter
(S15/S16) -- everything that does not need the GPU.  Mirrors the reference's own
test style (tests/test_cli.py there: run the CLI, assert on exit status)."""
import io
import os
import subprocess
import sys

import numpy as np

from hymet_b200 import cli
from hymet_b200.tsv import fmt_g, screen_lines

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = os.path.join(ROOT, "bin", "mash")


def run(*args):
    return subprocess.run([sys.executable, MASH] + list(args), capture_output=True, text=True)


def test_usage_and_errors():
    r = run("screen")
    assert r.returncode == 0 and "mash screen [options] <queries>.msh <mixture>" in r.stdout
    r = run("screen", "-h")
    assert r.returncode == 0 and "-w" in r.stdout
    r = run("screen", "db.txt", "x.fna")
    assert r.returncode == 1 and "does not look like a sketch (.msh)" in r.stderr and r.stdout == ""
    r = run("screen", "-v", "2", "db.msh", "x.fna")
    assert r.returncode == 1
    r = run("screen", "-q", "db.msh", "x.fna")
    assert r.returncode == 1 and "ERROR" in r.stderr
    r = run("dist", "a", "b")
    assert r.returncode == 2
    r = run("--version")
    assert r.returncode == 0


def test_screen_without_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        return
    from hymet_b200 import msh as mshfmt
    db = mshfmt.SketchDB(k=21, s=10, names=["a"], comments=[""], lengths=np.array([5], np.uint64),
                         offsets=np.array([0, 2], np.uint64), hashes=np.array([1, 2], np.uint64))
    p = str(tmp_path / "d.msh")
    mshfmt.write_msh(p, db)
    fa = tmp_path / "q.fna"; fa.write_text(">a\nACGT\n")
    r = run("screen", p, str(fa))
    assert r.returncode == 1 and "ERROR" in r.stderr and r.stdout == ""   # never a silent CPU result


def test_percent_g_formatting():
    assert fmt_g(1.0) == "1" and fmt_g(0.0) == "0"
    assert fmt_g(0.99956958100674098) == "0.99957"
    assert fmt_g(0.71968567300115205) == "0.719686"
    assert fmt_g(4.1498519816435698e-06) == "4.14985e-06"
    assert fmt_g(1e-05) == "1e-05" and fmt_g(0.0001) == "0.0001"
    assert fmt_g(1.3594475697585294e-23) == "1.35945e-23"


def test_reporting_rules_s15():
    shared = [0, 5, 1000, 1]
    sizes = [1000] * 4
    ident = [0.0, 0.777011, 1.0, 0.719686]
    pv = [1.0, 4.1e-6, 0.0, 0.95]
    names = ["GCF_%d" % i for i in range(4)]
    com = ["c%d" % i for i in range(4)]
    med = [0, 2, 7, 1]
    d = screen_lines(shared, sizes, med, ident, pv, names, com)                 # defaults: -i 0 -v 1
    assert [l.split("\t")[4] for l in d] == ["GCF_1", "GCF_2", "GCF_3"]
    assert d[1] == "1\t1000/1000\t7\t0\tGCF_2\tc2\n"
    d = screen_lines(shared, sizes, med, ident, pv, names, com, 0.0, 0.9)       # HYMET: -v 0.9
    assert [l.split("\t")[4] for l in d] == ["GCF_1", "GCF_2"]
    d = screen_lines(shared, sizes, med, ident, pv, names, com, -1.0, 1.0)      # -i -1 prints zero-hit sketches
    assert len(d) == 4 and d[0].startswith("0\t0/1000\t0\t1\tGCF_0")
    d = screen_lines(shared, sizes, med, ident, pv, names, com, 0.75, 1.0)
    assert [l.split("\t")[4] for l in d] == ["GCF_1", "GCF_2"]


def test_downstream_parsers_accept_our_lines():
    # scripts/limit_candidates.py:97-122 reads col 1 as float and col 5 as the name;
    # scripts/downloadDB.py:106-111 takes the first two '_' pieces of col 5.
    line = screen_lines([7], [1000], [1], [0.789], [1e-9], ["GCF_000005845.2_ASM584v2_genomic.fna"], ["x y"])[0]
    parts = line.rstrip("\n").split("\t")
    assert len(parts) == 6 and float(parts[0]) == 0.789
    assert "_".join(parts[4].split("_")[:2]) == "GCF_000005845.2"
