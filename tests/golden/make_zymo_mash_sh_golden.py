#!/usr/bin/env python
"""Golden outputs of the reference's UNMODIFIED scripts/mash.sh on the zymo fixture (run in the build
container: needs /root/reference; `mash` = the oracle CLI, `bc` = the stand-in of tests/test_stage_cpu.py).
Committed as zymo_mash_sh.json so that hymet_b200.stage.select is pinned on boxes without the reference."""
import base64
import gzip
import json
import os
import stat
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import _oracle as orc  # noqa: E402
from tests.test_stage_cpu import FAKE_BC  # noqa: E402

orc.build()
out = {}
for thr in ("0.9", "0.82"):
    d = tempfile.mkdtemp()
    os.makedirs(os.path.join(d, "input")); os.makedirs(os.path.join(d, "bin")); os.makedirs(os.path.join(d, "out"))
    open(os.path.join(d, "input", "sample_0.fna"), "wb").write(gzip.open(os.path.join(HERE, "zymo_query.fna.gz"), "rb").read())
    os.symlink(orc.BIN, os.path.join(d, "bin", "mash"))
    bc = os.path.join(d, "bin", "bc")
    open(bc, "w").write(FAKE_BC)
    os.chmod(bc, os.stat(bc).st_mode | stat.S_IXUSR)
    files = [os.path.join(d, "out", f) for f in ("screen.tab", "filtered.tab", "sorted.tab", "top_hits.tab", "selected.txt")]
    env = dict(os.environ, LC_ALL="C", PATH=os.path.join(d, "bin") + os.pathsep + os.environ["PATH"])
    p = subprocess.run(["bash", "/root/reference/scripts/mash.sh", os.path.join(d, "input"), os.path.join(HERE, "zymo25.msh")] + files + [thr],
                       capture_output=True, env=env, check=True)
    got = [open(f, "rb").read() for f in files]
    assert got[0] == open(os.path.join(HERE, "zymo_screen.tsv"), "rb").read()
    out[thr] = {"filtered": got[1].decode(), "sorted": got[2].decode(), "top_hits": got[3].decode(), "selected": got[4].decode(),
                "log": p.stdout.decode()}
    print(thr, got[4].count(b"\n"), "selected;", p.stdout.decode().splitlines()[-3])
json.dump(out, open(os.path.join(HERE, "zymo_mash_sh.json"), "w"), indent=1)
