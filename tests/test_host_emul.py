"""Device arithmetic + host packer, emulated on the CPU, against the oracle.

csrc/kmer_core.cuh is host+device code; this compiles it with g++ together with
the product's FASTA packer and walks the packed words the way one CUDA thread per
word does.  It is a pre-GPU check of layout, rolling k-mers, canonical choice and
the MurmurHash3 arithmetic -- not a product path.
"""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import py_micro_oracle as po
from tests import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(ROOT, "tests", "host_emul", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostemul.so")
    srcs = [os.path.join(ROOT, "tests", "host_emul", "host_emul.cpp"),
            os.path.join(ROOT, "hymet_b200", "csrc", "fasta_pack.cpp")]
    deps = srcs + [os.path.join(ROOT, "hymet_b200", "csrc", "kmer_core.cuh"),
                   os.path.join(ROOT, "hymet_b200", "csrc", "fasta_pack.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so] + srcs + ["-lz"], check=True)
    L = C.CDLL(so)
    L.emul_pack_and_hash.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint32,
                                     C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64)]
    L.emul_pack_and_hash.restype = C.c_int64
    L.emul_bucket_of.argtypes = [C.c_uint64, C.c_uint32]; L.emul_bucket_of.restype = C.c_uint32
    L.emul_pair_reverse.argtypes = [C.c_uint64]; L.emul_pair_reverse.restype = C.c_uint64
    L.emul_split.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    L.emul_split.restype = C.c_int
    return L


def emul_hashes(L, text: bytes, k: int, seed: int = 42):
    out = np.zeros(len(text) + 8, np.uint64)
    st = (C.c_uint64 * 3)()
    n = L.emul_pack_and_hash(text, len(text), k, seed, out.ctypes.data_as(C.POINTER(C.c_uint64)), len(out), st)
    assert n >= 0
    return out[:n], tuple(int(x) for x in st)


def oracle_hashes(text: str, k: int, seed: int = 42):
    hs = []
    recs = po.parse_fasta(text)
    for _, s in recs:
        h, v = orc.hash_sequence(s.encode(), k, seed)
        hs.append(h[v])
    return (np.concatenate(hs) if hs else np.zeros(0, np.uint64)), recs


def rand_fasta(rng, n_rec, max_len, width=60, junk=0.02):
    lines = []
    for r in range(n_rec):
        L = rng.randrange(0, max_len)
        s = "".join(rng.choice("ACGTacgtNnRYKM-") if rng.random() < junk else rng.choice("ACGT") for _ in range(L))
        lines.append(">rec%d desc" % r)
        w = rng.choice([width, 70, 80, 10 ** 9])
        lines += [s[i:i + w] for i in range(0, len(s), w)]
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("k", [5, 11, 15, 16, 17, 20, 21, 24, 31, 32])
def test_emulated_thread_walk_equals_oracle(emul, k):
    rng = random.Random(100 + k)
    text = rand_fasta(rng, 12, 900)
    got, st = emul_hashes(emul, text.encode(), k)
    want, recs = oracle_hashes(text, k)
    assert st[0] == len(recs) and st[1] == sum(len(s) for _, s in recs)
    assert st[2] == st[1] + st[0]          # one separator position per record
    assert np.array_equal(got, want)       # same hashes, same stream order


def test_edge_inputs(emul):
    k = 21
    cases = [
        "",                                     # empty
        ">only_header\n",                       # record without sequence
        ">short\nACGT\n",                       # shorter than k (S6)
        ">exact\n" + "ACGTTGCAACGTTGCAACGTA\n",  # exactly k
        ">a\n" + "A" * 64 + "\n>b\n" + "C" * 64 + "\n",   # word-aligned records must not bridge
        ">a\n" + "ACGT" * 16 + "N" + "ACGT" * 16 + "\n",
        ">crlf\r\nACGTACGTACGTACGTACGTACGT\r\nACGTACGT\r\n",
        ">nonl\nACGTACGTACGTACGTACGTACGTACG",   # no trailing newline
        ">x\nACGTACGTACGT ACGTACGTACGTACGTACGTACGT\n",     # blank inside a line is a non-alphabet byte
        "junk before header\n>x\n" + "GATTACA" * 9 + "\n\n\n>y\n\n" + "TTGACCA" * 9 + "\n",
        "@q1\n" + "ACGTTGCA" * 6 + "\n+\n" + "I" * 48 + "\n@q2\n" + "GGATCCAA" * 6 + "\n+q2\n" + "@" * 48 + "\n",
    ]
    for text in cases:
        got, st = emul_hashes(emul, text.encode(), k)
        if text.startswith("@"):
            want = np.concatenate([orc.hash_sequence(("ACGTTGCA" * 6).encode(), k)[0],
                                   orc.hash_sequence(("GGATCCAA" * 6).encode(), k)[0]])
            assert st[0] == 2
        else:
            want, _ = oracle_hashes(text, k)
        assert np.array_equal(got, want), text[:40]


def test_emulated_matches_oracle_parser_on_screen_golden(emul, golden_dir):
    import json
    g = json.load(open(os.path.join(golden_dir, "screen_small.json")))
    got, _ = emul_hashes(emul, g["fasta"].encode(), g["k"])
    want, _ = oracle_hashes(g["fasta"], g["k"])
    assert np.array_equal(got, want)
    # and the C oracle's own parser agrees on the k-mer count
    hs = [np.array([int(h) for h in r["hashes"]], np.uint64) for r in g["db"]]
    offsets = np.concatenate([[0], np.cumsum([len(h) for h in hs])]).astype(np.uint64)
    db = orc.OracleDB.from_arrays(g["k"], g["s"], 42, offsets, np.concatenate(hs),
                                  np.array([r["length"] for r in g["db"]], np.uint64))
    assert db.screen_text(g["fasta"].encode()).n_kmers == len(got)


def test_pair_reverse_and_bucket_range(emul):
    rng = random.Random(5)
    for _ in range(200):
        x = rng.getrandbits(64)
        want = 0
        for p in range(32):
            want |= ((x >> (2 * p)) & 3) << (2 * (31 - p))
        assert emul.emul_pair_reverse(x) == want
    for nb in (1, 7, 1000, 2 ** 20 + 3, 2 ** 31 + 11):
        bs = [emul.emul_bucket_of(rng.getrandbits(40), nb) for _ in range(2000)]
        assert max(bs) < nb
        if nb >= 1000:   # bottom-s style small keys still spread over the table
            assert len(set(b * 16 // nb for b in bs)) == 16


def test_split_records_preserves_kmers(emul):
    rng = random.Random(77)
    text = rand_fasta(rng, 40, 1500).encode()
    bounds = (C.c_uint64 * 128)()
    n = emul.emul_split(text, len(text), 7, 1, bounds, 64)
    assert 1 <= n <= 64
    spans = [(bounds[2 * i], bounds[2 * i + 1]) for i in range(n)]
    assert spans[0][0] == 0 and spans[-1][1] == len(text)
    assert all(spans[i][1] == spans[i + 1][0] for i in range(n - 1))
    assert all(text[b:b + 1] == b">" for b, _ in spans[1:])
    whole, _ = emul_hashes(emul, text, 21)
    parts = np.concatenate([emul_hashes(emul, text[b:e], 21)[0] for b, e in spans])
    assert np.array_equal(whole, parts)


@pytest.mark.parametrize("multiline,crlf", [(False, False), (True, False), (True, True)])
def test_split_fastq_records_preserves_kmers(emul, multiline, crlf):
    """FASTQ is splittable although '@' also starts quality lines: split_records cuts where the packer's own
    walk (record_starts) sees a header.  Packing the spans one by one gives the k-mers of the whole text."""
    rng = random.Random(5)
    recs = ["".join(rng.choice("ACGT") for _ in range(rng.choice([0, 1, 30, 75, 150, 400]))) for _ in range(300)]
    fq = make_fastq(recs, multiline, crlf, True).encode()
    bounds = (C.c_uint64 * 256)()
    for parts, min_span in ((7, 1), (40, 1), (3, 5000)):
        n = emul.emul_split(fq, len(fq), parts, min_span, bounds, 128)
        assert 2 <= n <= 128, n
        spans = [(bounds[2 * i], bounds[2 * i + 1]) for i in range(n)]
        assert spans[0][0] == 0 and spans[-1][1] == len(fq)
        assert all(spans[i][1] == spans[i + 1][0] for i in range(n - 1))
        assert all(fq[b:b + 1] == b"@" for b, _ in spans)
        whole, st = emul_hashes(emul, fq, 21)
        got = [emul_hashes(emul, fq[b:e], 21) for b, e in spans]
        assert np.array_equal(whole, np.concatenate([g[0] for g in got]))
        assert st[0] == sum(g[1][0] for g in got) == len(recs)


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_premultiplied_table_hash_equals_murmur(emul, k):
    """The streaming kernel takes the first multiply of every 64-bit murmur lane from a table of
    pre-multiplied four-letter words (kmer_core.cuh: hash_canonical_premul).  For every k: equal to
    the plain formulation and to the oracle's MurmurHash3 of the ASCII k-mer."""
    emul.emul_hash_both.argtypes = [C.c_uint64, C.c_int, C.c_uint32, C.POINTER(C.c_uint64)]
    emul.emul_hash_both.restype = None
    rng = random.Random(k)
    out = (C.c_uint64 * 2)()
    for t in range(300):
        cl = rng.getrandbits(64) if t > 3 else [0, (1 << 64) - 1, 0x5555555555555555, 0xAAAAAAAAAAAAAAAA][t]
        seed = 42 if t % 3 else rng.getrandbits(32)
        emul.emul_hash_both(cl, k, seed, out)
        assert out[0] == out[1], (k, hex(cl))
        if t < 40:
            kmer = bytes(b"ACGT"[(cl >> (2 * i)) & 3] for i in range(k))     # first base least significant
            h1 = orc.murmur(kmer, seed)[0]
            assert out[1] == (h1 if k > 16 else h1 & 0xFFFFFFFF)


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_windowed_msb_canonical_equals_rolling(emul, k):
    """k_stream's per-thread formulation (128-bit window + funnel shifts, MSB-first canonical k-mer,
    permuted pre-multiplied tables) gives the same 32 hashes per word as the rolling LSB-first one
    that the oracle comparison above pins."""
    emul.emul_window_vs_roll.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    emul.emul_window_vs_roll.restype = C.c_int
    rng = random.Random(100 + k)
    a, b = (C.c_uint64 * 32)(), (C.c_uint64 * 32)()
    specials = [0, (1 << 64) - 1, 0x5555555555555555, 0xAAAAAAAAAAAAAAAA, 0x1B1B1B1B1B1B1B1B, 0xE4E4E4E4E4E4E4E4]
    for t in range(200):
        prev = specials[t % 6] if t < 12 else rng.getrandbits(64)
        cur = specials[(t // 2) % 6] if t < 12 else rng.getrandbits(64)
        if t % 7 == 0:   # palindromic neighbourhoods: forward == reverse complement ties
            cur = prev
        assert emul.emul_window_vs_roll(prev, cur, k, 42, a, b) == 0, (k, hex(prev), hex(cur), list(a)[:4], list(b)[:4])


def make_fastq(records, multiline=False, crlf=False, final_newline=True):
    """FASTQ text whose quality strings contain every printable symbol, so quality lines start
    with '@', '>' and '+' now and then (the parser must count symbols, not look at line starts)."""
    out = []
    for i, s in enumerate(records):
        q = "".join(chr(33 + ((j * 7 + i * 13) % 94)) for j in range(len(s)))
        if multiline:
            sl = "\n".join(s[j:j + 61] for j in range(0, len(s), 61))
            ql = "\n".join(q[j:j + 61] for j in range(0, len(q), 61))
        else:
            sl, ql = s, q
        out.append("@read%d some text\n%s\n+%s\n%s\n" % (i, sl, "read%d" % i if i % 2 else "", ql))
    t = "".join(out)
    if not final_newline:
        t = t.rstrip("\n")
    if crlf:
        t = t.replace("\n", "\r\n")
    return t


@pytest.mark.parametrize("multiline,crlf,final_newline", [(False, False, True), (True, False, True), (False, True, True),
                                                          (True, True, False), (False, False, False)])
def test_fastq_records_pack_like_fasta(emul, multiline, crlf, final_newline):
    """Row a6 for FASTQ (kseq rules): header '@', sequence lines until '+', then as many quality
    symbols as the sequence had.  The product packer must produce exactly the k-mers of the sequences."""
    rng = random.Random(77)
    recs = []
    for r in range(40):
        L = rng.choice([0, 1, 20, 21, 22, 60, 61, 62, 200, 1000])
        recs.append("".join(rng.choice("ACGTacgtNn") if rng.random() < 0.03 else rng.choice("ACGT") for _ in range(L)))
    fq = make_fastq(recs, multiline, crlf, final_newline)
    assert any(l.startswith("@") and not l.startswith("@read") for l in fq.replace("\r", "").split("\n")) or not multiline
    for k in (16, 21, 31):
        got, st = emul_hashes(emul, fq.encode(), k)
        want = [orc.hash_sequence(s.encode(), k, 42) for s in recs if len(s)]
        want = np.concatenate([h[v] for h, v in want]) if want else np.zeros(0, np.uint64)
        assert got.tolist() == want.tolist(), (k, multiline, crlf)
        assert st[0] == len(recs) and st[1] == sum(len(s) for s in recs)
    # and the oracle's own parser reads the FASTQ as it reads the equivalent FASTA
    fa = "".join(">read%d\n%s\n" % (i, s) for i, s in enumerate(recs))
    a, b = orc.sketch_text(fq.encode(), 21, 300), orc.sketch_text(fa.encode(), 21, 300)
    assert a[0].tolist() == b[0].tolist() and a[1] == b[1]
