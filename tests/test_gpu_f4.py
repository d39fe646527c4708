"""SURVEY.md 8f rank 4 on the GPU: the weighted-LCA kernel against the reference's own output (golden) and,
bit for bit, against the oracle on random cases (confidence compared as doubles, not as printed text)."""
import os
import random

import numpy as np
import pytest

from hymet_b200 import lca
from oracle import lca_oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f4")


def test_classifier_output_equals_reference(tmp_path):
    out = str(tmp_path / "classified.tsv")
    done, total = lca.classify(os.path.join(G, "hits.paf"), os.path.join(G, "detailed_taxonomy.tsv"),
                               os.path.join(G, "taxonomy_hierarchy.tsv"), out)
    assert open(out, "rb").read() == open(os.path.join(G, "classified_reference.tsv"), "rb").read()
    assert (done, total) == (99, 121)


def test_kernel_is_bit_exact_with_the_oracle_on_random_votes():
    import ctypes as C

    from hymet_b200 import _abi
    rng = random.Random(11)
    _abi.init(0)
    n_tax, n_q = 300, 5000
    names = np.zeros((n_tax, 8), np.uint32)
    for t in range(n_tax):
        depth = rng.choice([0, 3, 5, 7, 8, 8])
        for r in range(depth):
            names[t, r] = 0 if rng.random() < 0.05 else 1 + (t * 7 + r) % rng.choice([2, 3, 5, 40])
    q_off, tax, w = [0], [], []
    for q in range(n_q):
        for _ in range(rng.choice([0, 1, 1, 2, 3, 6, 17, 40])):
            tax.append(-1 if rng.random() < 0.1 else rng.randrange(n_tax if q % 3 else 6))
            w.append(rng.choice([0.0, 1.0, 0.5, 1 / 3, rng.random(), rng.random() * rng.randrange(1, 50)]))
        q_off.append(len(tax))
    q_off = np.asarray(q_off, np.uint64); tax = np.asarray(tax, np.int32); w = np.asarray(w, np.float64)
    on = np.zeros((n_q, 8), np.uint32); od = np.zeros(n_q, np.uint32); oc = np.zeros(n_q, np.float64); oa = np.zeros(n_q, np.uint8)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    _abi.check(_abi.load().hs_lca_weighted(n_q, p(q_off, C.c_uint64), p(tax, C.c_int32), p(w, C.c_double), n_tax,
                                           p(names, C.c_uint32), p(on, C.c_uint32), p(od, C.c_uint32), p(oc, C.c_double),
                                           p(oa, C.c_uint8)))
    want = lca_oracle.weighted_lca(q_off.tolist(), tax.tolist(), w.tolist(), names.tolist())
    for q, (chosen, conf, any_hit) in enumerate(want):
        assert od[q] == len(chosen) and on[q, :len(chosen)].tolist() == chosen, q
        assert oc[q] == conf and bool(oa[q]) == any_hit, (q, oc[q], conf)      # doubles, bit for bit
    assert sum(1 for c, _, _ in want if c) > 2000
