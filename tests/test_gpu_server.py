"""The resident table server on a real GPU: served `mash screen` / stage output is byte-identical to the
in-process drop-in, the table is built once, a changed sketch file is rebuilt."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from hymet_b200 import msh as mshfmt
from hymet_b200 import synth
from tests.test_gpu_parity import build_db

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = [sys.executable, os.path.join(ROOT, "bin", "mash")]
SRV = [sys.executable, os.path.join(ROOT, "bin", "hymet-screen-server")]


def make_db(path, genomes, s=1000):
    offsets, hashes, lengths = build_db(genomes, 21, s)
    db = mshfmt.SketchDB(k=21, s=s, names=[synth.gcf_name(i) for i in range(len(genomes))], comments=["g%d" % i for i in range(len(genomes))],
                         lengths=lengths, offsets=offsets, hashes=hashes)
    mshfmt.write_msh(path, db)


def test_served_screen_is_byte_identical_and_table_is_resident(tmp_path):
    rng = np.random.default_rng(41)
    genomes = [synth.random_genome(rng, 40_000) for _ in range(40)]
    dbp = str(tmp_path / "db.msh")
    make_db(dbp, genomes)
    q1, q2 = str(tmp_path / "a.fna"), str(tmp_path / "b.fna")
    open(q1, "wb").write(synth.to_fasta(synth.cut_contigs(rng, genomes[:8], 400_000, 0.01, median=5000.0), "a"))
    open(q2, "wb").write(synth.to_fasta(synth.cut_contigs(rng, genomes[20:25], 200_000, 0.03, median=3000.0), "b"))
    env0 = dict(os.environ)
    env0.pop("HYMET_SCREEN_SERVER", None)
    sock = str(tmp_path / "run" / "gpu0.sock")
    env = dict(env0, HYMET_SCREEN_SERVER="1", HYMET_SCREEN_SOCKET=sock)
    want = {}
    for name, args in (("a", ["-p", "4", "-v", "0.9", dbp, q1]), ("w", ["-w", dbp, q1, q2]), ("rel", ["-p", "2", "db.msh", "b.fna"])):
        want[name] = subprocess.run(MASH + ["screen"] + args, capture_output=True, env=env0, cwd=str(tmp_path))
        assert want[name].returncode == 0 and want[name].stdout.count(b"\n") >= 3
    assert subprocess.run(SRV + ["start"], capture_output=True, env=env).returncode == 0
    try:
        for name, args in (("a", ["-p", "4", "-v", "0.9", dbp, q1]), ("w", ["-w", dbp, q1, q2]), ("rel", ["-p", "2", "db.msh", "b.fna"]),
                           ("a", ["-p", "4", "-v", "0.9", dbp, q1])):
            got = subprocess.run(MASH + ["screen"] + args, capture_output=True, env=env, cwd=str(tmp_path))
            assert got.returncode == 0 and got.stdout == want[name].stdout
            assert b"distinct hashes" in got.stderr
        st = subprocess.run(SRV + ["status"], capture_output=True, text=True, env=env)
        import json
        info = json.loads(st.stdout)
        assert info["served"] == 4 and len(info["tables"]) == 1          # one table for four requests
        # errors come back with mash's exit status, and do not take the daemon down
        bad = subprocess.run(MASH + ["screen", dbp, str(tmp_path / "missing.fna")], capture_output=True, env=env)
        assert bad.returncode == 1 and b"ERROR" in bad.stderr and bad.stdout == b""
        # a rewritten sketch file is a new table
        time.sleep(0.01)
        make_db(dbp, genomes[:30])
        again = subprocess.run(MASH + ["screen", "-p", "4", "-v", "0.9", dbp, q1], capture_output=True, env=env)
        fresh = subprocess.run(MASH + ["screen", "-p", "4", "-v", "0.9", dbp, q1], capture_output=True, env=env0)
        assert again.returncode == 0 and again.stdout == fresh.stdout
    finally:
        subprocess.run(SRV + ["stop"], capture_output=True, env=env)
    # with the daemon gone the same command line works in-process again
    back = subprocess.run(MASH + ["screen", "-p", "4", "-v", "0.9", dbp, q1], capture_output=True, env=env)
    assert back.returncode == 0 and back.stdout == fresh.stdout
