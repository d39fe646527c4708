"""The oracle against the golden vectors (CPU only).

Pins oracle/mash_screen_oracle.c (and the pure-Python micro-oracle) to
third-party known answers: Appleby's MurmurHash3 build, mpmath p-values, and
each other.  See tests/golden/make_golden.py for provenance.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import py_micro_oracle as po
from tests import _oracle as orc


def load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_murmur_kat_c_and_python(golden_dir):
    for v in load(golden_dir, "murmur_kat.json"):
        data = bytes.fromhex(v["hex"])
        want = (int(v["h1"], 16), int(v["h2"], 16))
        assert orc.murmur(data, v["seed"]) == want
        assert po.murmur3_x64_128(data, v["seed"]) == want


def test_murmur_survey_appendix_c():
    # SURVEY.md Appendix C (mash semantics: seed 42, h1 kept; low 32 bits for k <= 16)
    assert orc.murmur(b"AGCTTTTCATTCTGACTGCAA")[0] == 0xFADEE799080B0BA2
    assert orc.murmur(b"TTGCAGTCAGAATGAAAAGCT")[0] == 0xE594EA5044D4CBC8
    assert orc.murmur(b"ACGTACGTACGTACGT")[0] & 0xFFFFFFFF == 0xAC055887
    assert orc.murmur(b"foo", 0)[0] == 0xE271865701F54561


def test_use64_boundary():
    assert orc.lib().orc_use64(16) == 0 and orc.lib().orc_use64(17) == 1
    assert not po.use64(16) and po.use64(17)


@pytest.mark.parametrize("k", [11, 16, 17, 21, 31, 32])
def test_hash_sequence_matches_python(k):
    rng = random.Random(k)
    seq = "".join(rng.choice("ACGTacgtNnRY") if rng.random() < 0.05 else rng.choice("ACGT") for _ in range(700))
    h, v = orc.hash_sequence(seq.encode(), k)
    want = dict(po.kmer_hashes(seq, k))
    assert int(v.sum()) == len(want)
    for i in range(len(h)):
        if v[i]:
            assert int(h[i]) == want[i]
        else:
            assert i not in want


def test_canonical_strand_symmetry():
    rng = random.Random(3)
    seq = "".join(rng.choice("ACGT") for _ in range(500))
    rc = seq.translate(str.maketrans("ACGT", "TGCA"))[::-1]
    h1, _ = orc.hash_sequence(seq.encode(), 21)
    h2, _ = orc.hash_sequence(rc.encode(), 21)
    assert np.array_equal(h1, h2[::-1])
    # palindromic 20-mer + k=20: fwd == rc, tie -> forward
    pal = "ACGTACGTACGTACGTACGT"
    assert orc.hash_sequence(pal.encode(), 20)[0][0] == po.murmur3_x64_128(pal.encode())[0]


def test_pvalue_identity_kat(golden_dir):
    for v in load(golden_dir, "pvalue_kat.json"):
        p = orc.lib().orc_pvalue(v["x"], v["set_size"], 4.0 ** v["k"], v["n"])
        want = float(v["p"])
        if want == 0.0 or want < 1e-300:
            assert p == pytest.approx(want, rel=1e-9, abs=1e-320)
        else:
            assert abs(p - want) <= 1e-13 * want, v
        ident = orc.lib().orc_identity(v["x"], v["n"], v["k"])
        assert abs(ident - float(v["identity"])) <= 4e-16, v
    # formatting (S16) of the Appendix C rows
    assert orc.fmt_g(orc.lib().orc_identity(991, 1000, 21)) == "0.99957"
    assert orc.fmt_g(orc.lib().orc_pvalue(5, 10 ** 9, 4.0 ** 21, 1000)) == "4.14985e-06"
    assert orc.fmt_g(orc.lib().orc_pvalue(30, 10 ** 10, 4.0 ** 21, 1000)) == "1.35945e-23"
    assert orc.lib().orc_pvalue(0, 10 ** 9, 4.0 ** 21, 1000) == 1.0


def test_pvalue_python_exact_agrees():
    for x, n, ss, k in [(1, 50, 10 ** 7, 21), (3, 64, 5 * 10 ** 8, 21), (10, 40, 10 ** 6, 16)]:
        a = orc.lib().orc_pvalue(x, ss, 4.0 ** k, n)
        b = po.pvalue(x, ss, k, n)
        assert abs(a - b) <= 1e-13 * b


def _golden_db(g):
    hs = [np.array([int(h) for h in r["hashes"]], np.uint64) for r in g["db"]]
    offsets = np.concatenate([[0], np.cumsum([len(h) for h in hs])]).astype(np.uint64)
    lengths = np.array([r["length"] for r in g["db"]], np.uint64)
    return hs, offsets, np.concatenate(hs), lengths


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("mode", ["plain", "wta"])
def test_screen_small_golden(golden_dir, mode, threads):
    g = load(golden_dir, "screen_small.json")
    hs, offsets, hashes, lengths = _golden_db(g)
    db = orc.OracleDB.from_arrays(g["k"], g["s"], g["seed"], offsets, hashes, lengths)
    r = db.screen_text(g["fasta"].encode(), threads=threads, wta=(mode == "wta"))
    want = g["results"][mode]
    assert r.shared.tolist() == want["shared"]
    assert r.median.tolist() == want["median"]
    assert r.set_size == want["set_size"]
    assert [str(int(h)) for h in r.mixture] == want["mixture"]
    np.testing.assert_allclose(r.identity, want["identity"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(r.pvalue, want["pvalue"], rtol=1e-12, atol=0)


def test_sketch_matches_python(golden_dir):
    g = load(golden_dir, "screen_small.json")
    recs = po.parse_fasta(g["fasta"])
    h, total = orc.sketch_text(g["fasta"].encode(), 21, 50)
    assert total == sum(len(s) for _, s in recs)
    assert h.tolist() == po.sketch([s for _, s in recs], 21, 50)


def test_counts_wrap_and_multiplicity():
    # S8: duplicates count, both strands collapse
    rng = random.Random(9)
    g = "".join(rng.choice("ACGT") for _ in range(300))
    rc = g.translate(str.maketrans("ACGT", "TGCA"))[::-1]
    sk = po.sketch([g], 21, 1000)
    db = orc.OracleDB.from_arrays(21, 1000, 42, [0, len(sk)], sk, [300])
    fa = (">a\n%s\n>b\n%s\n>c\n%s\n" % (g, rc, g)).encode()
    r = db.screen_text(fa)
    assert int(r.shared[0]) == len(sk) and int(r.median[0]) == 3 and r.identity[0] == 1.0
    assert set(r.counts_per_entry.tolist()) == {3}


def test_zymo_real_genome_golden_is_reproduced_by_the_oracle(golden_dir):
    """tests/golden/zymo_*: 25 real RefSeq assemblies from the reference's own case study, sketched
    and screened by tests/golden/make_zymo_golden.py.  The oracle must keep reproducing the committed
    TSVs from the committed sketch file and query (the GPU test compares the CUDA path with the same files)."""
    import subprocess
    orc.build()
    msh, q = os.path.join(golden_dir, "zymo25.msh"), os.path.join(golden_dir, "zymo_query.fna.gz")
    for extra, name in (([], "zymo_screen.tsv"), (["-w"], "zymo_screen_w.tsv")):
        r = subprocess.run([orc.BIN, "screen", "-p", "3", "-v", "0.9"] + extra + [msh, q], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        want = open(os.path.join(golden_dir, name), "rb").read()
        assert r.stdout == want and want.count(b"\n") >= 10
    db = orc.OracleDB.load_msh(msh)
    assert db.n_refs == 25 and db.k == 21 and db.s == 1000
    assert db.name(0).startswith("GCF_") and db.name(0).endswith("_genomic.fna")
