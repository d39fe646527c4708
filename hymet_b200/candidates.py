"""Candidate limiting after the Mash stage (first half).

Mirror of /root/reference/scripts/limit_candidates.py (called at run_hymet_cami.sh:122): the same
command line, the same output file, the same summary line.  What it computes
(limit_candidates.py:97-122, 189-232):

  score(name)   = the best column-1 value of that name (column 5) over all `--score-file`s -- the
                  `mash screen` TSVs this repo's drop-in writes; unknown names score -inf
  species key   = with --dedupe, the species taxid of the name's assembly accession ("GCF_x.y" = the
                  first two '_'-separated pieces) from NCBI's assembly_summary files, the accession
                  itself when unknown; without --dedupe every name is its own key
  result        = names by descending score (ties: input order), first name of every key, at most --max

Host-side text work by nature (a few thousand lines); nothing here touches the GPU.  The assembly
summaries are read if present; like the reference with --no-download, a missing summary just means
"no species information" (there is no network in the deployment this was built for: the download the
reference attempts is not reproduced, and absence is handled exactly as its failure path is).
"""
from __future__ import annotations

import argparse
import csv
import os
import sys
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

SUMMARY_FILES = ("assembly_summary_refseq.txt", "assembly_summary_genbank.txt")   # limit_candidates.py:28-31, in this order
DEFAULT_MAX = 5000


def read_names(path: str) -> List[str]:
    with open(path, "r", encoding="utf-8") as fh:
        return [ln.strip() for ln in fh if ln.strip()]


def best_scores(paths: Iterable[str]) -> Dict[str, float]:
    """limit_candidates.py:97-122: column 1 as float, column 5 as name, keep the maximum per name;
    short lines, empty names, unparsable scores and unreadable files are skipped silently."""
    best: Dict[str, float] = {}
    for p in paths:
        if not os.path.exists(p):
            continue
        try:
            with open(p, "r", encoding="utf-8", errors="ignore") as fh:
                for line in fh:
                    if not line.strip():
                        continue
                    f = line.rstrip("\n").split("\t")
                    if len(f) < 5:
                        continue
                    name = f[4].strip()
                    if not name:
                        continue
                    try:
                        sc = float(f[0])
                    except ValueError:
                        continue
                    if name not in best or sc > best[name]:
                        best[name] = sc
        except OSError:
            continue
    return best


def species_map(directory: str) -> Dict[str, Tuple[str, str]]:
    """accession -> (species taxid or accession, organism name or accession) from whichever assembly
    summaries exist in `directory` (limit_candidates.py:163-186; later files overwrite earlier ones)."""
    out: Dict[str, Tuple[str, str]] = {}
    for name in SUMMARY_FILES:
        p = os.path.join(directory, name)
        if not os.path.exists(p):
            continue
        try:
            with open(p, "r", encoding="utf-8", errors="ignore") as fh:
                for row in csv.reader(fh, delimiter="\t"):
                    if not row or row[0].startswith("#") or len(row) < 8:
                        continue
                    acc = row[0].strip()
                    if not acc:
                        continue
                    taxid = (row[6] or row[5]).strip()
                    out[acc] = (taxid or acc, row[7].strip() or acc)
        except OSError:
            continue
    return out


def accession_of(name: str) -> str:
    """"GCF_000005845.2_ASM584v2_genomic.fna" -> "GCF_000005845.2" (limit_candidates.py:189-193)."""
    p = name.split("_", 2)
    return p[0] + "_" + p[1] if len(p) >= 2 else name


def limit(names: Sequence[str], scores: Dict[str, float], species: Dict[str, Tuple[str, str]], dedupe: bool,
          max_keep: int) -> Tuple[List[str], int]:
    """-> (kept names in output order, number of distinct keys seen among them)."""
    ninf = float("-inf")
    order = sorted(range(len(names)), key=lambda i: (-scores.get(names[i], ninf), i))
    kept: List[str] = []
    seen = set()
    for i in order:
        nm = names[i]
        if dedupe:
            acc = accession_of(nm)
            key = species.get(acc, (acc, acc))[0]
        else:
            key = nm
        if key in seen:
            continue
        seen.add(key)
        kept.append(nm)
        if max_keep > 0 and len(kept) >= max_keep:
            break
    return kept, len(seen)


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="Limit Mash candidate genomes with optional species-level deduplication.")
    ap.add_argument("--selected", required=True)
    ap.add_argument("--output", required=True)
    ap.add_argument("--score-file", action="append", default=[], dest="score_files")
    ap.add_argument("--assembly-dir", default=None)
    ap.add_argument("--max", type=int, default=DEFAULT_MAX)
    ap.add_argument("--dedupe", action="store_true")
    ap.add_argument("--log", default=None)
    ap.add_argument("--no-download", action="store_true",
                    help="accepted for compatibility: this implementation never downloads")
    a = ap.parse_args(argv)
    if a.max <= 0:
        raise SystemExit("The --max value must be greater than zero.")
    names = read_names(a.selected)
    if not names:
        raise SystemExit("No candidates found in %s" % a.selected)
    scores = best_scores(a.score_files)
    adir = a.assembly_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data",
                                          "downloaded_genomes", "assembly_summaries")
    species = {}
    if a.dedupe:
        os.makedirs(adir, exist_ok=True)
        species = species_map(adir)
    kept, n_keys = limit(names, scores, species, a.dedupe, a.max)
    tmp = a.output + ".tmp"
    with open(tmp, "w", encoding="utf-8") as fh:
        for nm in kept:
            fh.write(nm + "\n")
    os.replace(tmp, a.output)
    summary = ("[limit_candidates] kept %d / %d candidates (%d unique keys) %s"
               % (len(kept), len(names), n_keys if a.dedupe else len(kept), "(species dedupe)" if a.dedupe else ""))
    print(summary)
    if a.log:
        d = os.path.dirname(a.log)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(a.log, "a", encoding="utf-8") as fh:
            fh.write(summary.rstrip("\n") + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
