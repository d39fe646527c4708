"""Deterministic synthetic genomes, contig sets and sketch databases.

The mutation model is the one HYMET's own test-data script uses
(/root/reference/testdataset/mutationGCF.py:4-18): every A/C/G/T is replaced, with
probability ``rate``, by a uniformly chosen *different* base; anything else is kept.
Contig lengths follow a log-normal fitted to the reference's Zymo assembly
(/root/reference/case/truth/zymo_mc/zymo_mc_vs_refs.paf column 2: median ~15 kb,
p90 ~56 kb, min 1 kb, cap 6.5 Mb).

NumPy only (host); bench.py has the torch/GPU versions of the same generators for the
Gbp-scale workloads.  Bases are handled as codes A=0 C=1 G=2 T=3 (4 = 'N').
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)


def random_genome(rng: np.random.Generator, n: int, n_frac: float = 0.0) -> np.ndarray:
    g = rng.integers(0, 4, size=n, dtype=np.uint8)
    if n_frac > 0:
        n_runs = max(1, int(n * n_frac / 50))
        for st in rng.integers(0, max(1, n - 100), size=n_runs):
            g[st:st + int(rng.integers(1, 100))] = 4
    return g


def mutate(codes: np.ndarray, rate: float, rng: np.random.Generator) -> np.ndarray:
    """mutationGCF.py:4-18 on base codes."""
    out = codes.copy()
    hit = (rng.random(len(codes)) < rate) & (codes < 4)
    out[hit] = (codes[hit] + 1 + rng.integers(0, 3, size=int(hit.sum()), dtype=np.uint8)) % 4
    return out


def revcomp(codes: np.ndarray) -> np.ndarray:
    r = codes[::-1].copy()
    m = r < 4
    r[m] = 3 - r[m]
    return r


def contig_lengths(rng: np.random.Generator, total: int, median: float = 15000.0, sigma: float = 1.0,
                   lo: int = 1000, hi: int = 6_500_000) -> List[int]:
    out, acc = [], 0
    while acc < total:
        L = int(min(hi, max(lo, rng.lognormal(np.log(median), sigma))))
        L = min(L, total - acc) if total - acc >= lo else total - acc
        out.append(L)
        acc += L
    return out


def cut_contigs(rng: np.random.Generator, genomes: Sequence[np.ndarray], total: int, rate: float,
                rc_frac: float = 0.5, **kw) -> List[np.ndarray]:
    contigs = []
    for L in contig_lengths(rng, total, **kw):
        g = genomes[int(rng.integers(0, len(genomes)))]
        L = min(L, len(g))
        st = int(rng.integers(0, len(g) - L + 1))
        c = mutate(g[st:st + L], rate, rng) if rate > 0 else g[st:st + L].copy()
        if rng.random() < rc_frac:
            c = revcomp(c)
        contigs.append(c)
    return contigs


def to_fasta(records: Sequence[np.ndarray], prefix: str = "contig", width: int = 80, lower_frac: float = 0.0,
             rng: np.random.Generator = None) -> bytes:
    parts = []
    for i, c in enumerate(records):
        s = ASCII[c].tobytes()
        if lower_frac and rng is not None and rng.random() < lower_frac:
            s = s.lower()
        parts.append(b">%s_%d len=%d\n" % (prefix.encode(), i, len(c)))
        if width:
            parts.append(b"\n".join(s[j:j + width] for j in range(0, len(s), width)))
        else:
            parts.append(s)
        parts.append(b"\n")
    return b"".join(parts)


def decoy_sketches(rng: np.random.Generator, n: int, s: int, g_lo: float = 1.5e6, g_hi: float = 8e6,
                   bits: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    """Sketches of unrelated genomes without the genomes: the s smallest of G uniform
    `bits`-bit values are the partial sums of exponential gaps of mean 2^bits/G.
    Returns (hashes [n, s] ascending uint64, genome lengths [n])."""
    G = rng.uniform(g_lo, g_hi, size=n)
    gaps = rng.exponential(1.0, size=(n, s)) * ((2.0 ** bits) / G)[:, None]
    h = np.cumsum(np.maximum(gaps, 1.0), axis=1)
    h = np.minimum(h, 2.0 ** bits - 2.0 ** (bits - 52)).astype(np.uint64)
    # strictly ascending after the float->int cast
    h += np.arange(s, dtype=np.uint64)[None, :]
    return h, G.astype(np.uint64)


def gcf_name(i: int, tag: str = "synth") -> str:
    """File-name shaped like RefSeq's, as HYMET's downstream parsing expects
    (scripts/downloadDB.py:106-111, scripts/limit_candidates.py:188-192)."""
    return "GCF_%09d.1_%s%d_genomic.fna" % (i + 1, tag, i)
