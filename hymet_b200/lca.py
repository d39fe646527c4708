"""HYMET's contig classifier with the weighted-LCA vote on the GPU (second half).

Mirror of /root/reference/scripts/classification_cami.py (called at run_hymet_cami.sh:175): same
command line (--paf --taxonomy --hierarchy --output --processes), same output TSV
(Query, Lineage, Taxonomic Level, Confidence with four decimals), same reading rules for the
two taxonomy tables and the PAF.  What is where:

  host   the text: taxonomy / hierarchy tables (classification_cami.py:65-177), PAF lines
         (:185-207), the identifier lookup with its fallbacks (:211-249) -- one dictionary probe per
         DISTINCT target name -- and the output formatting (:333-340);
  GPU    the vote itself (:251-308): `hs_lca_weighted`, one thread per query, bit-exact with the
         reference's double arithmetic (csrc/lca_kernels.cu).  `--processes` is accepted and ignored:
         there is no process pool, every query is voted on at once.

No CPU implementation of the vote lives here: without a B200 the classifier fails loudly
(the checker's restatement is oracle/lca_oracle.py, test infrastructure).
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import gzip
import re
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import check

RANKS = ["superkingdom", "phylum", "class", "order", "family", "genus", "species", "strain"]
RANK_OF = {"domain": 0, "kingdom": 0, "sk": 0, "k": 0, "phylum": 1, "p": 1, "class": 2, "c": 2,
           "order": 3, "o": 3, "family": 4, "f": 4, "genus": 5, "g": 5, "species": 6, "s": 6, "subspecies": 7, "ss": 7,
           "strain": 7}     # as the reference's RANK_ALIAS (:17-27): the tag "superkingdom" itself is NOT among them
ASSEMBLY_RE = re.compile(r"GC[AF]_\d+(?:\.\d+)?(?:_PRJ[A-Z]+\d+)?")
CONTIG_RE = re.compile(r"(NC_\d+\.\d+|NZ_[A-Z]{2}\d+\.\d+|NZ_[A-Z]{5}\d+\.\d+|CP\d+\.\d+|CM\d+\.\d+|[A-Z]{2}_\d+\.\d+)")
_SPLIT_IDS = re.compile(r"[;|,\s]+")
_SPLIT_LINEAGE = re.compile(r"[;|]+")
_SPLIT_HEAD = re.compile(r"[|\s]+")

csv.field_size_limit(1024 * 1024 * 1024)   # the Identifiers column can be enormous


def _remember(m: Dict[str, str], token: str, taxid: str) -> None:
    """First writer wins, for the token and for its version-less form (classification_cami.py:44-54)."""
    token = (token or "").strip()
    if not token:
        return
    m.setdefault(token, taxid)
    if "." in token:
        m.setdefault(token.split(".", 1)[0], taxid)


def load_taxonomy(path: str) -> Dict[str, str]:
    """identifier -> TaxID from detailed_taxonomy.tsv (classification_cami.py:65-105): per row, in this
    order, every GCF/GCA accession found in any column, every token of the Identifiers column, then
    every contig-style accession found in Identifiers or any column."""
    m: Dict[str, str] = {}
    with open(path, "r", newline="") as fh:
        rd = csv.DictReader(fh, delimiter="\t")
        if "TaxID" not in (rd.fieldnames or []):
            raise RuntimeError("TaxID column not found in taxonomy file")
        for row in rd:
            taxid = (row.get("TaxID") or "").strip()
            if not taxid:
                continue
            for v in row.values():
                if v:
                    for acc in ASSEMBLY_RE.findall(v):
                        _remember(m, acc, taxid)
            ids = row.get("Identifiers") or ""
            for tok in _SPLIT_IDS.split(ids) if ids else ():
                if tok.strip():
                    _remember(m, tok, taxid)
            for v in (ids,) + tuple(row.get(k) or "" for k in row.keys()):
                if v:
                    for acc in CONTIG_RE.findall(v):
                        _remember(m, acc, taxid)
    return m


def lineage_names(raw: str) -> List[str]:
    """One name per rank from the lineage encodings the reference accepts (classification_cami.py:107-160):
    `rank:name; ...`, `k__Bacteria; p__...`, or plain names in rank order (NA skipped)."""
    out = [""] * len(RANKS)
    s = (raw or "").strip()
    if not s:
        return out
    for sep in (":", "__"):
        if sep in s:
            for part in _SPLIT_LINEAGE.split(s):
                part = part.strip()
                if not part or sep not in part:
                    continue
                tag, name = part.split(sep, 1)
                r = RANK_OF.get(tag.strip().lower())
                name = name.strip()
                if r is not None and name:
                    out[r] = name
            return out
    plain = [p.strip() for p in _SPLIT_LINEAGE.split(s) if p.strip() and p.strip().upper() != "NA"]
    for i, name in enumerate(plain[:len(RANKS)]):
        out[i] = name
    return out


def load_hierarchy(path: str) -> Dict[str, List[str]]:
    """TaxID -> names by rank from taxonomy_hierarchy.tsv (classification_cami.py:162-177; a later row of
    the same TaxID replaces the earlier one)."""
    h: Dict[str, List[str]] = {}
    with open(path, "r", newline="") as fh:
        rd = csv.DictReader(fh, delimiter="\t")
        if "TaxID" not in (rd.fieldnames or []) or "Lineage" not in (rd.fieldnames or []):
            raise RuntimeError("Hierarchy file must have TaxID and Lineage columns")
        for row in rd:
            tid = (row.get("TaxID") or "").strip()
            if tid:
                h[tid] = lineage_names((row.get("Lineage") or "").strip())
    return h


def parse_paf(path: str):
    """-> (queries in order of first appearance, per query [(target, coverage)], alignments per target)
    (classification_cami.py:185-207: columns 1, 2, 6 and 11; lines with fewer than 11 columns or starting
    with '#' are skipped; unparsable lengths count as coverage 0)."""
    order: List[str] = []
    per_q: Dict[str, List[Tuple[str, float]]] = {}
    per_t: Dict[str, int] = {}
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")) as fh:
        for line in fh:
            if not line or line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 11:
                continue
            try:
                qlen, block = int(f[1]), int(f[10])
            except Exception:                       # noqa: BLE001  (the reference swallows everything here)
                qlen, block = 0, 0
            cov = (block / qlen) if qlen > 0 else 0.0
            q = f[0]
            if q not in per_q:
                per_q[q] = []
                order.append(q)
            per_q[q].append((f[5], cov))
            per_t[f[5]] = per_t.get(f[5], 0) + 1
    return order, per_q, per_t


def lookup_keys(target: str) -> List[str]:
    """Keys tried against the identifier map, in order (classification_cami.py:211-238): the name, its
    version-less form, its first token, every embedded assembly / contig accession -- each followed by
    its version-less form."""
    keys: List[str] = []

    def add(x: str) -> None:
        if x and x not in keys:
            keys.append(x)
        if x and "." in x:
            v = x.split(".", 1)[0]
            if v not in keys:
                keys.append(v)

    add(target)
    add(_SPLIT_HEAD.split(target)[0])
    for g in ASSEMBLY_RE.findall(target):
        add(g)
    for a in CONTIG_RE.findall(target):
        add(a)
    return keys


def taxid_of(target: str, tax: Dict[str, str]) -> Optional[str]:
    for k in lookup_keys(target):
        t = tax.get(k)
        if t:
            return t
    return None


class Encoded:
    """The vote's input as the kernel takes it: integer-coded, CSR over the queries."""

    def __init__(self, q_off, tax_rows, weights, names, name_of):
        self.q_off, self.tax_rows, self.weights, self.names, self.name_of = q_off, tax_rows, weights, names, name_of


def encode(order: Sequence[str], per_q, per_t: Dict[str, int], tax: Dict[str, str], hier: Dict[str, List[str]]) -> Encoded:
    """Host half of the vote: resolve every DISTINCT target name to a taxid once (classification_cami.py:240-249),
    intern the lineage names per rank, weigh every alignment (coverage x alignments of its target, :298)."""
    tax_row: Dict[str, int] = {}
    rows: List[List[int]] = []
    name_id: List[Dict[str, int]] = [dict() for _ in RANKS]
    name_of: List[List[str]] = [[""] for _ in RANKS]
    target_row: Dict[str, int] = {}

    def row_of_target(t: str) -> int:
        r = target_row.get(t)
        if r is None:
            tid = taxid_of(t, tax)
            if tid is None:
                r = -1
            else:
                r = tax_row.get(tid)
                if r is None:
                    names = hier.get(tid) or [""] * len(RANKS)     # no lineage: a taxid that never has a name
                    ids = []
                    for k, nm in enumerate(names[:len(RANKS)]):
                        if not nm:
                            ids.append(0)
                            continue
                        i = name_id[k].get(nm)
                        if i is None:
                            i = len(name_of[k])
                            name_id[k][nm] = i
                            name_of[k].append(nm)
                        ids.append(i)
                    ids += [0] * (len(RANKS) - len(ids))
                    r = len(rows)
                    rows.append(ids)
                    tax_row[tid] = r
            target_row[t] = r
        return r

    q_off = np.zeros(len(order) + 1, np.uint64)
    tax_rows: List[int] = []
    weights: List[float] = []
    for i, q in enumerate(order):
        for t, cov in per_q[q]:
            tax_rows.append(row_of_target(t))
            weights.append(cov * per_t.get(t, 1))
        q_off[i + 1] = len(tax_rows)
    names = np.asarray(rows, np.uint32).reshape(-1, len(RANKS)) if rows else np.zeros((0, len(RANKS)), np.uint32)
    return Encoded(q_off, np.asarray(tax_rows, np.int32), np.asarray(weights, np.float64), names, name_of)


def decode(enc: Encoded, out_names, out_depth, out_conf) -> List[Tuple[str, str, float]]:
    """Kernel output -> (lineage string, level, confidence) per query (classification_cami.py:286-288)."""
    res = []
    for i in range(len(enc.q_off) - 1):
        d = int(out_depth[i])
        if d == 0:
            res.append(("Unknown", "root", 0.0))
            continue
        lin = "; ".join("%s:%s" % (RANKS[k], enc.name_of[k][int(out_names[i][k])]) for k in range(d))
        res.append((lin, RANKS[d - 1], float(out_conf[i])))
    return res


def vote(order: Sequence[str], per_q, per_t: Dict[str, int], tax: Dict[str, str], hier: Dict[str, List[str]], device: int = 0):
    """encode -> hs_lca_weighted on the GPU -> decode: per query (lineage string, level, confidence), in `order`."""
    _abi.init(device)
    enc = encode(order, per_q, per_t, tax, hier)
    n_q = len(order)
    out_names = np.zeros((max(n_q, 1), len(RANKS)), np.uint32)
    out_depth = np.zeros(max(n_q, 1), np.uint32)
    out_conf = np.zeros(max(n_q, 1), np.float64)
    out_any = np.zeros(max(n_q, 1), np.uint8)
    p = lambda arr, t: arr.ctypes.data_as(C.POINTER(t))
    n_a, n_t = len(enc.tax_rows), len(enc.names)
    check(_abi.load().hs_lca_weighted(n_q, p(enc.q_off, C.c_uint64), p(enc.tax_rows, C.c_int32) if n_a else None,
                                      p(enc.weights, C.c_double) if n_a else None, n_t,
                                      p(enc.names, C.c_uint32) if n_t else None, p(out_names, C.c_uint32),
                                      p(out_depth, C.c_uint32), p(out_conf, C.c_double), p(out_any, C.c_uint8)))
    return decode(enc, out_names, out_depth, out_conf)


def classify(paf: str, taxonomy: str, hierarchy: str, output: str, device: int = 0) -> Tuple[int, int]:
    """classification_cami.py:312-343: -> (classified, total)."""
    tax = load_taxonomy(taxonomy)
    hier = load_hierarchy(hierarchy)
    order, per_q, per_t = parse_paf(paf)
    res = vote(order, per_q, per_t, tax, hier, device)
    with open(output, "w", newline="") as fh:
        wr = csv.writer(fh, delimiter="\t")
        wr.writerow(["Query", "Lineage", "Taxonomic Level", "Confidence"])
        for q, (lin, lvl, conf) in zip(order, res):
            wr.writerow([q, lin, lvl, "%.4f" % conf])
    return sum(1 for lin, _, _ in res if lin != "Unknown"), len(res)


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="HYMET taxonomy classifier (robust identifiers + weighted LCA on the GPU)")
    ap.add_argument("--paf", required=True)
    ap.add_argument("--taxonomy", required=True)
    ap.add_argument("--hierarchy", required=True)
    ap.add_argument("--output", required=True)
    ap.add_argument("--processes", type=int, default=4, help="accepted for compatibility; the vote runs on the GPU")
    a = ap.parse_args(argv)
    import os
    try:
        done, total = classify(a.paf, a.taxonomy, a.hierarchy, a.output, int(os.environ.get("HYMET_SCREEN_DEVICE", "0")))
    except _abi.HsError as e:
        sys.stderr.write("ERROR: %s\n" % e.msg)
        return 1
    sys.stderr.write("Classification complete. Results saved to %s\nClassified: %d/%d (%.1f%%)\n"
                     % (a.output, done, total, 100.0 * done / total if total else 0.0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
