"""Locate and drive a REAL `mash` binary (marbl/Mash), when one exists on the box.

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  The reference's arithmetic for this path is the
third-party `mash` that scripts/mash.sh:14 forks (/root/reference/run_hymet_cami.sh:72 only checks
`command -v mash`; environment.yml:9 installs it unpinned).  It is absent from the build image, so
parity is pinned to the oracle restatement -- but the moment a real binary shows up (PATH,
baseline/_ref/, or $HYMET_REAL_MASH) bench.py times it instead of the oracle and byte-compares its
TSV with the CUDA path's, and tests/test_real_mash.py stops skipping.  Nothing under hymet_b200/
imports this module.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import time
from typing import List, Optional, Tuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUR_SHIM = os.path.realpath(os.path.join(ROOT, "bin", "mash"))


def _is_real(path: str) -> bool:
    """A candidate counts as real mash if it is executable, is not this repo's shim (or a link to
    it), is not a script, and answers `--version` with a version number."""
    try:
        if not (os.path.isfile(path) and os.access(path, os.X_OK)):
            return False
        if os.path.realpath(path) == OUR_SHIM:
            return False
        with open(path, "rb") as fh:
            if fh.read(2) == b"#!":
                return False            # the repo's drop-in (or the oracle CLI wrapped in a script), not marbl/Mash
        r = subprocess.run([path, "--version"], capture_output=True, text=True, timeout=20)
        v = (r.stdout + r.stderr).strip()
        return r.returncode == 0 and v[:1].isdigit() and "hymet" not in v.lower()
    except Exception:
        return False


def find_mash() -> Optional[str]:
    cands: List[str] = []
    if os.environ.get("HYMET_REAL_MASH"):
        cands.append(os.environ["HYMET_REAL_MASH"])
    ref = os.path.join(ROOT, "baseline", "_ref")
    cands += [os.path.join(ref, "bin", "mash"), os.path.join(ref, "mash")]
    for d in os.environ.get("PATH", "").split(os.pathsep):
        if d:
            cands.append(os.path.join(d, "mash"))
    w = shutil.which("mash")
    if w:
        cands.append(w)
    for c in cands:
        if _is_real(c):
            return c
    return None


def version(mash: str) -> str:
    r = subprocess.run([mash, "--version"], capture_output=True, text=True, timeout=20)
    return (r.stdout + r.stderr).strip()


def screen(mash: str, msh: str, inputs: List[str], threads: int, extra: Tuple[str, ...] = ()) -> Tuple[bytes, float, str]:
    """`mash screen -p threads [extra] msh inputs...` -> (TSV bytes, wall seconds, stderr)."""
    cmd = [mash, "screen", "-p", str(threads), *extra, msh, *inputs]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("real mash failed (%d): %s" % (r.returncode, r.stderr.decode("utf-8", "replace")[-500:]))
    return r.stdout, dt, r.stderr.decode("utf-8", "replace")


def sketch(mash: str, fasta: str, out_prefix: str, k: int = 21, s: int = 1000) -> str:
    subprocess.run([mash, "sketch", "-k", str(k), "-s", str(s), "-o", out_prefix, fasta], check=True, capture_output=True)
    return out_prefix + ".msh"


def info_table(mash: str, msh: str) -> str:
    """`mash info -t`: one line per reference (hashes, length, id, comment)."""
    return subprocess.run([mash, "info", "-t", msh], check=True, capture_output=True, text=True).stdout
