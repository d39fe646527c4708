"""oracle/py_micro_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, independent restatement of the `mash screen` rules (SURVEY.md
Appendix A, S1-S18) in pure Python, for tiny inputs only.  It exists so that
the C oracle (oracle/mash_screen_oracle.c) and the CUDA path are both checked
against something that shares no code with either.  The path it restates is
the one HYMET invokes at /root/reference/scripts/mash.sh:14; the arithmetic
lives in the third-party `mash` binary (unpinned, environment.yml:9), so like
the C oracle this is **parity unpinned by the reference** and anchored on the
published algorithm + the known-answer vectors in tests/golden/.

Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

M64 = (1 << 64) - 1


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & M64


def _fmix(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & M64
    k ^= k >> 33
    return k


def murmur3_x64_128(data: bytes, seed: int = 42) -> Tuple[int, int]:
    """S2. Returns (h1, h2); Mash keeps h1 (or its low 32 bits, S1)."""
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    h1 = h2 = seed & 0xFFFFFFFF
    n = len(data)
    nb = n // 16
    for i in range(nb):
        k1 = int.from_bytes(data[16 * i:16 * i + 8], "little")
        k2 = int.from_bytes(data[16 * i + 8:16 * i + 16], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
        h1 = _rotl(h1, 27); h1 = (h1 + h2) & M64; h1 = (h1 * 5 + 0x52DCE729) & M64
        k2 = (k2 * c2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
        h2 = _rotl(h2, 31); h2 = (h2 + h1) & M64; h2 = (h2 * 5 + 0x38495AB5) & M64
    tail = data[16 * nb:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & M64; h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & M64; h1 ^= k1
    h1 ^= n; h2 ^= n
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    h1 = _fmix(h1); h2 = _fmix(h2)
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    return h1, h2


def use64(k: int) -> bool:
    return 4.0 ** k > 2.0 ** 32


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def kmer_hashes(seq: str, k: int, seed: int = 42) -> List[Tuple[int, int]]:
    """(start, hash) for every valid window of one record (S3-S5)."""
    s = seq.upper()
    out = []
    w64 = use64(k)
    for i in range(len(s) - k + 1):
        f = s[i:i + k]
        if any(c not in _COMP for c in f):
            continue
        r = "".join(_COMP[c] for c in reversed(f))
        h = murmur3_x64_128(min(f, r).encode(), seed)[0]
        out.append((i, h if w64 else h & 0xFFFFFFFF))
    return out


def parse_fasta(text: str) -> List[Tuple[str, str]]:
    recs, name, buf = [], None, []
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(buf)))
            name, buf = line[1:], []
        elif name is not None:
            buf.append(line)  # kseq keeps every byte of a sequence line
    if name is not None:
        recs.append((name, "".join(buf)))
    return recs


def sketch(records: Iterable[str], k: int, s: int, seed: int = 42) -> List[int]:
    """`mash sketch`: s smallest distinct hashes over all records of a genome."""
    hs = set()
    for r in records:
        if len(r) >= k:
            hs.update(h for _, h in kmer_hashes(r, k, seed))
    return sorted(hs)[:s]


def identity(shared: int, size: int, k: int) -> float:
    if shared == size:
        return 1.0
    if shared == 0:
        return 0.0
    return (shared / size) ** (1.0 / k)


def pvalue(x: int, set_size: int, k: int, size: int) -> float:
    """S14 by direct summation of the binomial upper tail (math.comb, exact ints)."""
    if x == 0:
        return 1.0
    r = 1.0 / (1.0 + (4.0 ** k) / set_size)
    from fractions import Fraction
    rf = Fraction(r)
    tot = Fraction(0)
    for j in range(x, size + 1):
        tot += math.comb(size, j) * rf ** j * (1 - rf) ** (size - j)
    return float(tot)


def screen(db: Sequence[Tuple[Sequence[int], int]], records: Iterable[str], k: int, s: int,
           seed: int = 42, wta: bool = False) -> Dict[str, list]:
    """db = [(sorted hashes, genome length)].  Returns per-reference columns (S7-S18)."""
    table: Dict[int, List[int]] = {}
    for i, (hs, _) in enumerate(db):
        for h in hs:
            table.setdefault(h, []).append(i)
    counts = {h: 0 for h in table}
    mix = set()
    for r in records:
        if len(r) < k:
            continue
        for _, h in kmer_hashes(r, k, seed):
            mix.add(h)
            if h in counts:
                counts[h] = (counts[h] + 1) & 0xFFFFFFFF
    bottom = sorted(mix)[:s]
    W = 64 if use64(k) else 32
    set_size = int((2.0 ** W) * len(bottom) / float(bottom[-1])) if bottom else 0
    n = len(db)
    depths: List[List[int]] = [[] for _ in range(n)]
    for h, c in counts.items():
        if c:
            for i in table[h]:
                depths[i].append(c)
    if wta:
        score = [identity(len(depths[i]), len(db[i][0]), k) for i in range(n)]
        depths = [[] for _ in range(n)]
        for h, c in counts.items():
            if not c:
                continue
            best = None
            for i in table[h]:  # ascending index; full ties -> highest index (documented rule)
                if best is None or score[i] > score[best] or (
                        score[i] == score[best] and db[i][1] >= db[best][1]):
                    best = i
            depths[best].append(c)
    out = dict(shared=[], median=[], identity=[], pvalue=[], set_size=set_size, mixture=bottom)
    for i in range(n):
        d = sorted(depths[i])
        sh = len(d)
        out["shared"].append(sh)
        out["median"].append(d[sh // 2] if sh else 0)
        out["identity"].append(identity(sh, len(db[i][0]), k))
        out["pvalue"].append(pvalue(sh, set_size, k, len(db[i][0])) if set_size else (1.0 if sh == 0 else 0.0))
    return out
