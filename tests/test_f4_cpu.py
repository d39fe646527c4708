"""SURVEY.md 8f rank 4 on the CPU: candidate limiting (pure host work) byte-for-byte against what the
reference's own scripts/limit_candidates.py wrote, and the classifier's host half + the oracle's vote
against what scripts/classification_cami.py wrote (tests/golden/f4/, made by tests/golden/make_lca_golden.py
by importing the reference).  The CUDA vote itself is checked in tests/test_gpu_f4.py."""
import csv
import io
import os

import numpy as np
import pytest

from hymet_b200 import candidates, lca
from oracle import lca_oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f4")


@pytest.mark.parametrize("tag,extra", [("plain", []), ("dedupe", ["--dedupe", "--no-download"]), ("cap", ["--max", "7"]),
                                       ("dedupe_cap", ["--dedupe", "--no-download", "--max", "4"])])
def test_limit_candidates_equals_reference(tmp_path, capsys, tag, extra):
    out, log = str(tmp_path / "out.txt"), str(tmp_path / "sub" / "limit.log")
    rc = candidates.main(["--selected", os.path.join(G, "selected_genomes.txt"), "--output", out,
                          "--score-file", os.path.join(G, "screen_a.tab"), "--score-file", os.path.join(G, "screen_b.tab"),
                          "--score-file", os.path.join(G, "missing.tab"), "--assembly-dir", os.path.join(G, "assembly_summaries"),
                          "--log", log] + extra)
    assert rc == 0
    assert open(out, "rb").read() == open(os.path.join(G, "limited_%s_reference.txt" % tag), "rb").read()
    want_log = open(os.path.join(G, "limit_%s_reference.log" % tag)).read()
    assert open(log).read() == want_log and capsys.readouterr().out == want_log


def test_limit_candidates_errors(tmp_path):
    empty = tmp_path / "empty.txt"
    empty.write_text("\n\n")
    with pytest.raises(SystemExit, match="No candidates found"):
        candidates.main(["--selected", str(empty), "--output", str(tmp_path / "o")])
    with pytest.raises(SystemExit, match="greater than zero"):
        candidates.main(["--selected", os.path.join(G, "selected_genomes.txt"), "--output", str(tmp_path / "o"), "--max", "0"])
    assert candidates.accession_of("GCF_000005845.2_ASM584v2_genomic.fna") == "GCF_000005845.2"
    assert candidates.accession_of("plain") == "plain"


def test_lineage_encodings_and_lookup_keys():
    assert lca.lineage_names("k__Bacteria; p__Firmicutes; g__Bacillus")[:2] == ["Bacteria", "Firmicutes"]
    assert lca.lineage_names("k__Bacteria; p__Firmicutes; g__Bacillus")[5] == "Bacillus"
    # the reference's alias table has no "superkingdom" entry: that tag is dropped
    assert lca.lineage_names("superkingdom:Bacteria; phylum:P; strain:S") == ["", "P", "", "", "", "", "", "S"]
    assert lca.lineage_names("Bacteria|P|NA|O") == ["Bacteria", "P", "O", "", "", "", "", ""]
    assert lca.lineage_names("") == [""] * 8
    assert lca.lookup_keys("lcl|NZ_CP005001.2 text") == ["lcl|NZ_CP005001.2 text", "lcl|NZ_CP005001", "lcl", "NZ_CP005001.2",
                                                          "NZ_CP005001"]
    keys = lca.lookup_keys("GCF_000000101.2_ASM1v1_genomic")
    assert keys[:2] == ["GCF_000000101.2_ASM1v1_genomic", "GCF_000000101"] and keys[2:4] == ["GCF_000000101.2", "GCF_000000101"][:1] + keys[3:4]
    assert "GCF_000000101.2" in keys and keys[-2:] == ["CF_000000101.2", "CF_000000101"]   # the reference's contig pattern also bites inside "GCF_"


def test_classifier_host_half_plus_oracle_vote_equals_reference():
    """encode (host) -> the ORACLE's vote -> decode (host) reproduces the reference's TSV byte for byte: pins the
    host logic and the oracle; the CUDA kernel is then compared with both on the GPU box."""
    tax = lca.load_taxonomy(os.path.join(G, "detailed_taxonomy.tsv"))
    hier = lca.load_hierarchy(os.path.join(G, "taxonomy_hierarchy.tsv"))
    order, per_q, per_t = lca.parse_paf(os.path.join(G, "hits.paf"))
    assert len(order) == 121 and "1017" not in hier and tax["NC_009000"] == "1000"
    enc = lca.encode(order, per_q, per_t, tax, hier)
    got = lca_oracle.weighted_lca(enc.q_off.tolist(), enc.tax_rows.tolist(), enc.weights.tolist(), enc.names.tolist())
    names = np.zeros((len(order), 8), np.uint32)
    depth = np.zeros(len(order), np.uint32)
    conf = np.zeros(len(order), np.float64)
    for i, (chosen, c, _) in enumerate(got):
        names[i, :len(chosen)] = chosen
        depth[i], conf[i] = len(chosen), c
    res = lca.decode(enc, names, depth, conf)
    buf = io.StringIO(newline="")
    wr = csv.writer(buf, delimiter="\t")
    wr.writerow(["Query", "Lineage", "Taxonomic Level", "Confidence"])
    for q, (lin, lvl, c) in zip(order, res):
        wr.writerow([q, lin, lvl, "%.4f" % c])
    assert buf.getvalue().encode() == open(os.path.join(G, "classified_reference.tsv"), "rb").read()
    assert sum(1 for lin, _, _ in res if lin == "Unknown") == 22
