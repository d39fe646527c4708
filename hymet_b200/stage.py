"""The whole Mash stage of HYMET in one process and ONE pass over the contigs.

Rows a2 + a3 of SURVEY.md 8a (8f rank 2).  The reference runs scripts/mash.sh three times
(/root/reference/run_hymet_cami.sh:85-97): each run is `mash screen` against one of
data/sketch1-3.msh (mash.sh:14), then sort / sort / a bc+awk threshold loop / cut
(mash.sh:15-55), and the three selected-genome lists are concatenated and `sort -u`ed.
Here the three sketch files become one table (hs_db_from_msh_multi), the contigs are streamed
once, and the per-file TSVs and every derived file are written byte-identically to what the
three mash.sh runs leave behind (C/POSIX collation, which is what the reference's container
uses: it sets no locale).

    bin/hymet-mash-stage [-p N] [--merge] INPUT_DIR THRESHOLD \
        DB.msh SCREEN_TAB FILTERED SORTED TOP_HITS SELECTED   [DB.msh ... SELECTED]...

The five outputs per sketch file are mash.sh's arguments 3-7, THRESHOLD its argument 8.
--merge appends the later SELECTED lists to the first one and sorts it uniquely
(run_hymet_cami.sh:91,96,97).  stdout carries mash.sh's own log lines.
"""
from __future__ import annotations

import fnmatch
import os
import re
import sys
from decimal import Decimal
from typing import Dict, List, Optional, Sequence, Tuple

_BLANK = b" \t"
_NUM = re.compile(rb"[ \t\n\v\f\r]*([+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|inf(?:inity)?|nan))", re.I)


# ---------------------------------------------------------------- sort(1) / awk / cut restated
def sort_field(line: bytes, f: int) -> bytes:
    """Key `-k f,f` of GNU sort without -t/-b: a field is its leading blanks plus the
    non-blanks that follow."""
    i, n = 0, len(line)
    for _ in range(f - 1):
        while i < n and line[i] in _BLANK:
            i += 1
        while i < n and line[i] not in _BLANK:
            i += 1
    b = i
    while i < n and line[i] in _BLANK:
        i += 1
    while i < n and line[i] not in _BLANK:
        i += 1
    return line[b:i]


def sort_u_k5(lines: Sequence[bytes]) -> List[bytes]:
    """`sort -u -k5,5` (mash.sh:15): order by field 5, keep the first line of every run of equal
    keys (-u makes the sort stable and drops the last-resort comparison)."""
    out: List[bytes] = []
    last = None
    for key, ln in sorted(((sort_field(l, 5), l) for l in lines), key=lambda t: t[0]):
        if key != last:
            out.append(ln)
            last = key
    return out


def _general_numeric(line: bytes) -> Tuple[int, float]:
    m = _NUM.match(line)
    if not m:
        return (0, 0.0)                    # no number: sorts below everything
    v = float(m.group(1).decode())
    if v != v:
        return (1, 0.0)                    # NaN: above "no number", below -inf
    return (2, v)


def sort_gr(lines: Sequence[bytes]) -> List[bytes]:
    """`sort -gr` (mash.sh:16): general-numeric value of the line's leading number, descending;
    equal values fall back to the whole line, bytewise, also reversed."""
    return sorted(lines, key=lambda l: (_general_numeric(l), l), reverse=True)


def awk_gt(lines: Sequence[bytes], threshold: str) -> List[bytes]:
    """`awk -v t=T '$1 > t'`: both sides look numeric, so awk compares doubles."""
    t = float(threshold)
    out = []
    for l in lines:
        f = l.split()
        if f:
            try:
                if float(f[0]) > t:
                    out.append(l)
                continue
            except ValueError:
                pass
        # non-numeric first field: awk falls back to string comparison
        if (f[0] if f else b"") > threshold.encode():
            out.append(l)
    return out


def cut_f5(lines: Sequence[bytes]) -> List[bytes]:
    """`cut -f5`: TAB-delimited field 5; a line without any TAB is printed whole."""
    out = []
    for l in lines:
        if b"\t" not in l:
            out.append(l)
            continue
        p = l.split(b"\t")
        out.append(p[4] if len(p) > 4 else b"")
    return out


def _bc(d: Decimal) -> str:
    """How bc prints a number: no leading zero before the point."""
    if d == 0:
        return "0"
    s = format(d, "f")
    if s.startswith("0."):
        return s[1:]
    if s.startswith("-0."):
        return "-" + s[2:]
    return s


def min_candidates(n_fna: int) -> int:
    """mash.sh:20-21: round-half-up of 3.25 * files, at least 5."""
    return max(5, (13 * n_fna + 2) // 4)


def select(screen_tab: bytes, n_fna: int, initial_threshold: str) -> Dict[str, object]:
    """mash.sh:15-55 on the bytes `mash screen` printed.  Returns the four derived files, the
    threshold used and the script's log."""
    lines = screen_tab.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    filtered = sort_u_k5(lines)
    ordered = sort_gr(filtered)
    need = min_candidates(n_fna)
    log = ["====================================", "Number of input sequences: %d" % n_fna,
           "Minimum expected candidates: %d" % need, "===================================="]
    cur_s, cur = initial_threshold, Decimal(initial_threshold)
    best_s, found, count = initial_threshold, False, 0
    while cur >= Decimal("0.70"):
        count = len(awk_gt(ordered, cur_s))
        log += ["Testing threshold: %s" % cur_s, "Candidates found: %d" % count]
        if count >= need:
            best_s, found = cur_s, True
            break
        cur = cur - Decimal("0.02")
        cur_s = _bc(cur)
    if not found:
        best_s = "0.71"
        count = len(awk_gt(ordered, best_s))
        log.append("No suitable threshold found. Using 0.70.")
    top = awk_gt(ordered, best_s)
    sel = cut_f5(top)
    log += ["====================================", "Final threshold used: %s" % best_s,
            "Candidates found: %d" % count, "===================================="]
    j = lambda ls: b"".join(l + b"\n" for l in ls)
    return {"filtered": j(filtered), "sorted": j(ordered), "top_hits": j(top), "selected": j(sel),
            "threshold": best_s, "count": count, "log": "\n".join(log) + "\n"}


def merge_selected(parts: Sequence[bytes]) -> bytes:
    """run_hymet_cami.sh:91,96,97: cat the lists together, `sort -u` whole lines."""
    lines = set()
    for p in parts:
        ls = p.split(b"\n")
        if ls and ls[-1] == b"":
            ls.pop()
        lines.update(ls)
    return b"".join(l + b"\n" for l in sorted(lines))


def input_files(input_dir: str) -> Tuple[List[str], int]:
    """What `"$INPUT_DIR"/*.fna` expands to (no dot files, C order) and what
    `find -maxdepth 1 -name "*.fna" | wc -l` counts (dot files included)."""
    names = os.listdir(input_dir)
    n_find = sum(1 for n in names if fnmatch.fnmatchcase(n, "*.fna"))
    globbed = sorted((n for n in names if not n.startswith(".") and fnmatch.fnmatchcase(n, "*.fna")),
                     key=lambda n: n.encode())
    return [os.path.join(input_dir, n) for n in globbed], n_find


# ---------------------------------------------------------------- the fused stage
def run_stage(input_dir: str, threshold: str, jobs: Sequence[Sequence[str]], threads: int = 8, max_p: float = 0.9,
              merge: bool = False, device: int = 0, stdout=None, stderr=None, server=None) -> int:
    """jobs: (DB.msh, SCREEN_TAB, FILTERED, SORTED, TOP_HITS, SELECTED) per sketch file.
    `server`: the resident table server when this runs inside it (tables and screens outlive the call)."""
    stdout = stdout or sys.stdout
    stderr = stderr or sys.stderr
    from . import _lite as hs          # ctypes only: this is a one-shot process, numpy would be 15 % of it

    files, n_fna = input_files(input_dir)
    # What can go wrong with the DATA goes wrong the way it does in the reference: scripts/mash.sh has no
    # `set -e`, so a `mash screen` that fails (sketch file missing or malformed -- a deployment without the
    # custom sketch3.msh is ordinary --, no *.fna, no sequence records) prints its ERROR, leaves an EMPTY
    # screen.tab, and lines 15-55 still run on it: five empty files, the usual log, exit status 0
    # (run_hymet_cami.sh:90,95 add `|| true` and `cat ... 2>/dev/null || true` on top).  Only the GPU
    # path itself being unavailable is fatal here: no device, no library, a CUDA error.
    results: Dict[int, bytes] = {i: b"" for i in range(len(jobs))}
    n_gpus = max(1, int(os.environ.get("HYMET_SCREEN_GPUS", "1") or "1"))
    usable = []
    for i, j in enumerate(jobs):
        if not j[0].endswith(".msh"):
            stderr.write("ERROR: %s does not look like a sketch (.msh)\n" % j[0])
        else:
            usable.append(i)
    try:
        hs._abi.init(device)                   # no B200, no library: fail before anything is written (no CPU path)
        groups = []
        if usable:
            try:
                mg = None
                if server is not None:
                    ent = server.table([jobs[i][0] for i in usable], tolerate=True)
                    db = ent["db"]
                elif n_gpus > 1:                   # HYMET_SCREEN_GPUS: one table copy per GPU, inputs cut by bytes
                    mg = hs.MultiGpu([jobs[i][0] for i in usable], range(device, device + n_gpus), tolerate=True)
                    ent, db = None, mg.db
                else:
                    ent, db = None, hs.LiteDb([jobs[i][0] for i in usable], device, tolerate=True)
                for k_, msg in sorted(db.errors.items()):
                    stderr.write("ERROR: %s\n" % msg)
                if db.loaded:
                    groups = [(db, [usable[k_] for k_ in db.loaded], ent, mg)]
            except hs.HsError as e:
                if e.code in hs.DEVICE_ERRORS or "one at a time" not in e.msg:
                    raise
                # sketch files with different k / seed: one table each
                for i in usable:
                    try:
                        ent = server.table([jobs[i][0]]) if server is not None else None
                        groups.append((ent["db"] if ent else hs.LiteDb(jobs[i][0], device), [i], ent, None))
                    except hs.HsError as e2:
                        if e2.code in hs.DEVICE_ERRORS:
                            raise
                        stderr.write("ERROR: %s\n" % e2.msg)
        if not files:                          # the shell would hand mash the unexpanded pattern
            stderr.write("ERROR: could not open %s for reading.\n" % os.path.join(input_dir, "*.fna"))
            groups = []
        for db, idx, ent, mg in groups:
            try:
                if mg is not None:
                    scr = mg.screen(files, threads)
                else:
                    scr = server.screen_for(ent) if ent is not None else hs.LiteScreen(db)
                    for p in files:
                        scr.feed_fasta(p, threads)
                    scr.flush()
                if scr.stats()["n_records"] == 0:
                    stderr.write("ERROR: Did not find sequence records in inputs.\n")
                    continue
                for seg, (b, e) in zip(idx, db.segments):
                    results[seg] = "".join(scr.finish_lines(False, 0.0, max_p, b, e)).encode("utf-8", "surrogateescape")
            except hs.HsError as e:
                if e.code in hs.DEVICE_ERRORS:
                    raise
                stderr.write("ERROR: %s\n" % e.msg)
    except hs.HsError as e:
        stderr.write("ERROR: %s\n" % e.msg)
        return 1
    selected = []
    for i, j in enumerate(jobs):
        tab = results[i]
        r = select(tab, n_fna, threshold)
        for path, data in zip(j[1:6], (tab, r["filtered"], r["sorted"], r["top_hits"], r["selected"])):
            with open(path, "wb") as fh:
                fh.write(data)
        stdout.write(r["log"])
        selected.append(r["selected"])
    if merge and jobs:
        with open(jobs[0][5], "wb") as fh:
            fh.write(merge_selected(selected))
    stdout.flush()
    return 0


def _via_server(argv: List[str]) -> Optional[int]:
    """HYMET_SCREEN_SERVER=1|auto: let the resident table server run the stage (same files, same log)."""
    from . import server
    path = server.default_socket_path()
    req = {"op": "stage", "argv": list(argv), "cwd": os.getcwd()}
    try:
        return server.request(path, req)
    except OSError:
        pass
    if os.environ.get("HYMET_SCREEN_SERVER") != "auto":
        return None
    try:
        server.spawn_detached(path, int(os.environ.get("HYMET_SCREEN_DEVICE", "0")))
        return server.request(path, req)
    except OSError as e:
        sys.stderr.write("WARNING: no screen server (%s); running the stage in-process\n" % e)
        return None


def main(argv: Optional[List[str]] = None, stdout=None, stderr=None, cwd: Optional[str] = None, server=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    stdout = stdout or sys.stdout
    stderr = stderr or sys.stderr
    if server is None and os.environ.get("HYMET_SCREEN_SERVER", "0") not in ("", "0") and not any(a in ("-h", "--help") for a in argv):
        rc = _via_server(argv)
        if rc is not None:
            return rc
    threads, merge, max_p = 8, False, 0.9
    pos: List[str] = []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-h", "--help"):
            stdout.write(__doc__)
            return 0
        if a == "--merge":
            merge = True
        elif a in ("-p", "-v") and i + 1 < len(argv):
            try:
                if a == "-p":
                    threads = int(argv[i + 1])
                else:
                    max_p = float(argv[i + 1])
            except ValueError:
                stderr.write("ERROR: malformed numeric option value\n")
                return 1
            i += 1
        else:
            pos.append(a)
        i += 1
    if len(pos) < 8 or (len(pos) - 2) % 6:
        stderr.write(__doc__)
        return 2
    try:
        Decimal(pos[1])
    except Exception:
        stderr.write("ERROR: threshold must be a decimal number\n")
        return 1
    if cwd:      # served request: relative paths are the client's
        pos = [p if (i == 1 or os.path.isabs(p)) else os.path.join(cwd, p) for i, p in enumerate(pos)]
    jobs = [pos[2 + 6 * g: 8 + 6 * g] for g in range((len(pos) - 2) // 6)]
    return run_stage(pos[0], pos[1], jobs, threads, max_p, merge, int(os.environ.get("HYMET_SCREEN_DEVICE", "0")),
                     stdout=stdout, stderr=stderr, server=server)


if __name__ == "__main__":
    sys.exit(main())
