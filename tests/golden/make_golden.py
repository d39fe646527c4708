#!/usr/bin/env python
"""Generate the committed golden vectors in tests/golden/ (run once, here).

Nothing in the reference pins `mash screen` output (no .msh, no
screen.tab, tests/test_cli.py is dry-run only), so the vectors come from
*independent third parties present in this container*:

* murmur_kat.json   -- Austin Appleby's public-domain MurmurHash3.cpp as shipped
                       inside scikit-learn (sklearn/utils/src), compiled with g++
                       here; plus the SURVEY Appendix C known answers.
* pvalue_kat.json   -- mpmath (50 digits) regularised incomplete beta
                       I_r(x, n-x+1) == P[Binomial(n, r) >= x]  (S14), and
                       identity = (x/n)^(1/k) (S13) at 50 digits.
* screen_small.json -- a tiny end-to-end screen computed by the pure-Python
                       micro-oracle (oracle/py_micro_oracle.py): inputs + outputs.

The C oracle, the CUDA path and the C-ABI are all tested against these files.
Usage:  python tests/golden/make_golden.py
"""
import ctypes as C
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def appleby_lib():
    import sklearn
    src = os.path.join(os.path.dirname(sklearn.__file__), "utils", "src", "MurmurHash3.cpp")
    out = os.path.join(tempfile.mkdtemp(), "libappleby.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.MurmurHash3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_void_p]
    return L


def make_murmur():
    L = appleby_lib()
    rng = random.Random(20261018)
    vecs = []

    def add(data: bytes, seed: int):
        out = (C.c_uint64 * 2)()
        L.MurmurHash3_x64_128(data, len(data), seed, out)
        vecs.append(dict(hex=data.hex(), seed=seed, h1="%016x" % out[0], h2="%016x" % out[1]))

    for s in (b"foo", b"hello", b"", b"A" * 21, b"ACGTACGTACGTACGTACGTA", b"AGCTTTTCATTCTGACTGCAA",
              b"TTGCAGTCAGAATGAAAAGCT", b"ACGTACGTACGTACGTACGTACGTACGTACG", b"ACGTACGTACGTACGT"):
        for seed in (0, 42):
            add(s, seed)
    for n in range(0, 41):
        for _ in range(3):
            add(bytes(rng.choice(b"ACGT") for _ in range(n)), 42)
    for n in (1, 7, 15, 16, 17, 31, 32, 33, 64):
        add(bytes(rng.randrange(256) for _ in range(n)), rng.randrange(1 << 32))
    json.dump(vecs, open(os.path.join(HERE, "murmur_kat.json"), "w"), indent=0)
    print("murmur_kat.json", len(vecs))


def make_pvalue():
    import mpmath as mp
    mp.mp.dps = 60
    vecs = []
    cases = [(991, 1000, 3e9, 21), (5, 1000, 1e9, 21), (1, 1000, 1e7, 21), (2, 1000, 1e7, 21),
             (1, 400, 4.6e6, 21), (30, 1000, 1e10, 21), (1000, 1000, 1e9, 21), (999, 1000, 1e9, 21),
             (100, 1000, 5e9, 21), (500, 1000, 1e10, 21), (3, 5000, 1e9, 21), (50, 10000, 1e10, 21),
             (7000, 10000, 1e10, 21), (1, 1, 1e6, 21), (1, 1000, 1e9, 31), (10, 1000, 1e10, 31),
             (400, 1000, 1e10, 31), (1, 1000, 1e6, 16), (20, 1000, 1e6, 16), (200, 1000, 5e6, 16),
             (5, 1000, 3e6, 11), (700, 1000, 3e6, 11), (900, 1000, 3e6, 11), (1, 1000, 12345, 21),
             (2, 37, 98765432, 21), (36, 37, 98765432, 21), (12, 1000, 18446744073709551615, 21),
             (900, 1000, 18446744073709551615, 31)]
    rng = random.Random(7)
    for _ in range(60):
        n = rng.choice([100, 400, 1000, 5000, 10000])
        x = rng.randrange(1, n + 1)
        ss = int(10 ** rng.uniform(4, 11))
        k = rng.choice([16, 21, 31])
        cases.append((x, n, ss, k))
    for x, n, ss, k in cases:
        ss = int(ss)
        # r exactly as the double arithmetic of S14 produces it (inputs to I_r are then exact)
        r = 1.0 / (1.0 + (4.0 ** k) / float(ss))
        p = mp.betainc(x, n - x + 1, 0, mp.mpf(r), regularized=True)
        ident = mp.mpf(1) if x == n else mp.power(mp.mpf(float(x) / float(n)), mp.mpf(1.0 / k))
        vecs.append(dict(x=x, n=n, set_size=ss, k=k, r=repr(r), p=mp.nstr(p, 25), p_float=float(p),
                         identity=mp.nstr(ident, 25), identity_float=float(ident)))
    json.dump(vecs, open(os.path.join(HERE, "pvalue_kat.json"), "w"), indent=0)
    print("pvalue_kat.json", len(vecs))


def make_screen_small():
    from oracle import py_micro_oracle as po
    rng = random.Random(4242)
    k, s = 21, 64
    genomes = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(2500, 4000))) for _ in range(6)]
    # near-identical pair for -w, one genome with an N run, one shorter than s k-mers
    genomes.append(genomes[0][:1800] + "".join(rng.choice("ACGT") for _ in range(900)))
    genomes.append(genomes[1][:1200] + "NNNNNNNNNN" + genomes[1][1210:2600])
    genomes.append("".join(rng.choice("ACGT") for _ in range(60)))
    db = [(po.sketch([g], k, s), len(g)) for g in genomes]

    def mutate(seq, m):
        out = []
        for c in seq:
            if c in "ACGT" and rng.random() < m:
                out.append(rng.choice([b for b in "ACGT" if b != c]))
            else:
                out.append(c)
        return "".join(out)

    def rc(seq):
        return seq.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]

    contigs = [genomes[0][100:2100], rc(mutate(genomes[1][:2000], 0.01)), genomes[2][500:1500].lower(),
               genomes[2][700:1700], mutate(genomes[3], 0.05), genomes[6][1500:2600], "ACGTNNACGT",
               genomes[8], genomes[0][100:2100], "".join(rng.choice("ACGT") for _ in range(3000))]
    lines = []
    for i, c in enumerate(contigs):
        lines.append(">contig_%d some description" % i)
        for j in range(0, len(c), 70):
            lines.append(c[j:j + 70])
    fasta = "\n".join(lines) + "\n"
    res = {}
    for wta in (False, True):
        r = po.screen(db, contigs, k, s, 42, wta)
        res["wta" if wta else "plain"] = dict(shared=r["shared"], median=r["median"], identity=r["identity"],
                                             pvalue=r["pvalue"], set_size=r["set_size"],
                                             mixture=[str(h) for h in r["mixture"]])
    out = dict(k=k, s=s, seed=42, fasta=fasta,
               db=[dict(name="GCF_%09d.1_synth%d_genomic.fna" % (i + 1, i), length=ln,
                        hashes=[str(h) for h in hs]) for i, (hs, ln) in enumerate(db)],
               results=res)
    json.dump(out, open(os.path.join(HERE, "screen_small.json"), "w"), indent=0)
    print("screen_small.json refs", len(db), "shared", res["plain"]["shared"], "wta", res["wta"]["shared"])


if __name__ == "__main__":
    make_murmur()
    make_pvalue()
    make_screen_small()
