"""Parity against a REAL `mash` binary.  Skips when none is on the box (today's image has none: the
oracle restatement is then the only checker and parity is "unpinned").  When one appears
-- PATH, baseline/_ref/, or $HYMET_REAL_MASH -- these tests pin, with no further work:
  * the oracle (and the committed golden TSVs derived from it) against real mash on real genomes,
  * the .msh writer/reader (Appendix B ordinals) against `mash info` and `mash sketch`,
  * the CUDA CLI's bytes against real mash (GPU-marked).
"""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import real_mash

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = real_mash.find_mash()
needs_mash = pytest.mark.skipif(MASH is None, reason="no real mash binary on this box (PATH, baseline/_ref, $HYMET_REAL_MASH)")


def test_locator_ignores_the_repo_shim(tmp_path, monkeypatch):
    """bin/mash first on PATH is the deployment: it must never be mistaken for the real thing."""
    monkeypatch.setenv("PATH", os.path.join(ROOT, "bin") + os.pathsep + os.environ.get("PATH", ""))
    monkeypatch.delenv("HYMET_REAL_MASH", raising=False)
    found = real_mash.find_mash()
    assert found is None or os.path.realpath(found) != os.path.realpath(os.path.join(ROOT, "bin", "mash"))
    fake = tmp_path / "mash"
    fake.write_text("#!/bin/sh\necho 2.3\n")
    fake.chmod(0o755)
    monkeypatch.setenv("HYMET_REAL_MASH", str(fake))
    assert real_mash.find_mash() != str(fake)          # a script is not marbl/Mash


@needs_mash
@pytest.mark.parametrize("extra,name", [((), "zymo_screen.tsv"), (("-w",), "zymo_screen_w.tsv")])
def test_real_mash_reproduces_the_committed_goldens(tmp_path, golden_dir, extra, name):
    """The goldens were made by the oracle; real mash must print the same bytes (this is what pins the
    oracle).  Known interpretation risk: the -w full-tie order (S17) -- a mismatch confined to tied
    references is reported as such."""
    q = tmp_path / "q.fna"
    q.write_bytes(gzip.open(os.path.join(golden_dir, "zymo_query.fna.gz"), "rb").read())
    tsv, _, _ = real_mash.screen(MASH, os.path.join(golden_dir, "zymo25.msh"), [str(q)], 4, ("-v", "0.9") + extra)
    want = open(os.path.join(golden_dir, name), "rb").read()
    assert tsv == want, "real mash %s disagrees with the oracle-made golden %s" % (real_mash.version(MASH), name)


@needs_mash
def test_real_mash_reads_our_msh_and_we_read_its(tmp_path, golden_dir):
    from hymet_b200 import msh as mshfmt
    from tests import _oracle as orc
    # ours -> theirs
    db = mshfmt.read_msh(os.path.join(golden_dir, "zymo25.msh"))
    table = real_mash.info_table(MASH, os.path.join(golden_dir, "zymo25.msh")).splitlines()
    rows = [l.split("\t") for l in table if l and not l.startswith("#")]
    assert len(rows) == len(db.names)
    for i, r in enumerate(rows):
        assert int(r[0]) == int(db.offsets[i + 1] - db.offsets[i]) and int(r[1]) == int(db.lengths[i]) and r[2] == db.names[i]
    # theirs -> ours: sketch one genome with real mash, read it with the product's parser, compare with the oracle sketch
    text = gzip.open(os.path.join(golden_dir, "zymo_query.fna.gz"), "rb").read()
    fa = tmp_path / "g.fna"
    fa.write_bytes(text)
    p = real_mash.sketch(MASH, str(fa), str(tmp_path / "g"), 21, 1000)
    theirs = mshfmt.read_msh(p)
    ours, total = orc.sketch_text(text, 21, 1000)
    assert theirs.k == 21 and theirs.s == 1000 and theirs.seed == 42
    assert np.array_equal(theirs.hashes, ours) and int(theirs.lengths[0]) == total


@needs_mash
@pytest.mark.gpu
@pytest.mark.parametrize("extra", [(), ("-w",)])
def test_cuda_cli_bytes_equal_real_mash(tmp_path, golden_dir, extra):
    q = tmp_path / "q.fna"
    q.write_bytes(gzip.open(os.path.join(golden_dir, "zymo_query.fna.gz"), "rb").read())
    dbp = os.path.join(golden_dir, "zymo25.msh")
    want, _, _ = real_mash.screen(MASH, dbp, [str(q)], 4, ("-v", "0.9") + extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", "4", "-v", "0.9", *extra, dbp, str(q)],
                       capture_output=True)
    assert r.returncode == 0 and r.stdout == want
