#!/usr/bin/env python
"""Golden inputs and outputs for SURVEY.md 8f rank 4, made by IMPORTING THE REFERENCE in the build container
(needs /root/reference): scripts/classification_cami.py (main_process) and scripts/limit_candidates.py
(main) run on small synthetic tables written here; inputs and the reference's outputs are committed under
tests/golden/f4/ so that the restatements are pinned on boxes without the reference.

The case is built to hit the rules that matter: the three lineage encodings, the missing "superkingdom"
alias, version-less identifier fallbacks, targets without a taxid, taxids without a lineage, ties between
names (first maximum wins), weights that are not exactly representable, duplicate hits of one target
(abundance weighting), zero-length queries, candidates tied on score, species deduplication.
"""
import importlib.util
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "f4")
REF = "/root/reference/scripts"


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = random.Random(7)
    genera = ["Escherichia", "Salmonella", "Bacillus", "Listeria", "Pseudomonas", "Staphylococcus"]
    fam = {"Escherichia": "Enterobacteriaceae", "Salmonella": "Enterobacteriaceae", "Bacillus": "Bacillaceae",
           "Listeria": "Listeriaceae", "Pseudomonas": "Pseudomonadaceae", "Staphylococcus": "Staphylococcaceae"}
    tax_rows, hier_rows, targets = [], [], []
    for i in range(40):
        g = genera[i % len(genera)]
        tid = str(1000 + i)
        gcf = "GCF_%09d.%d" % (100 + i, 1 + i % 3)
        contig = "NZ_CP%06d.%d" % (5000 + i, 1 + i % 2) if i % 2 else "NC_%06d.1" % (9000 + i)
        ids = "; ".join([gcf, contig, "extra_%d" % i]) if i % 5 else "%s|%s,%s" % (gcf, contig, "x%d" % i)
        tax_rows.append((gcf + "_ASM%dv1" % i, tid, ids, "%s sp. %d" % (g, i)))
        sp = "%s species%d" % (g, i % 7)
        if i % 4 == 0:
            lin = "k__Bacteria; p__Phylum%d; c__Class%d; o__Order%d; f__%s; g__%s; s__%s" % (i % 2, i % 3, i % 3, fam[g], g, sp)
        elif i % 4 == 1:
            lin = "domain:Bacteria; phylum:Phylum%d; class:Class%d; order:Order%d; family:%s; genus:%s; species:%s; strain:str%d" % (
                i % 2, i % 3, i % 3, fam[g], g, sp, i)
        elif i % 4 == 2:
            lin = "superkingdom:Bacteria; phylum:Phylum%d; family:%s; genus:%s" % (i % 2, fam[g], g)   # first tag is ignored
        else:
            lin = "Bacteria|Phylum%d|NA|Order%d|%s|%s|%s" % (i % 2, i % 3, fam[g], g, sp)
        if i != 17:                       # taxid 1017 has no lineage at all
            hier_rows.append((tid, lin))
        targets.append((gcf, contig))
    with open(os.path.join(OUT, "detailed_taxonomy.tsv"), "w") as fh:
        fh.write("Assembly\tTaxID\tIdentifiers\tOrganism\n")
        for r in tax_rows:
            fh.write("\t".join(r) + "\n")
        fh.write("GCF_000000001.1_none\t\tNC_000001.1\tno taxid row\n")
    with open(os.path.join(OUT, "taxonomy_hierarchy.tsv"), "w") as fh:
        fh.write("TaxID\tLineage\n")
        for r in hier_rows:
            fh.write("\t".join(r) + "\n")
    with open(os.path.join(OUT, "hits.paf"), "w") as fh:
        fh.write("# comment line\n")
        for q in range(120):
            qlen = rng.choice([1000, 1537, 4096, 12345, 99991]) if q != 50 else 0
            n = rng.choice([1, 1, 2, 3, 5, 8, 13])
            for a in range(n):
                i = rng.randrange(40) if q % 11 else rng.choice([3, 3, 9, 15])       # q % 11 == 0: few targets, repeated
                gcf, contig = targets[i]
                style = rng.randrange(5)
                t = [contig, gcf + "_ASM%dv1_genomic" % i, contig.split(".")[0], "lcl|" + contig + " some text", "unknown_target_%d" % a][style]
                block = rng.randrange(1, max(2, qlen)) if qlen else 10
                cols = ["q%03d" % q, str(qlen), "0", str(block), "+", t, "5000000", "100", str(100 + block), str(block - 3), str(block), "60"]
                if q == 77 and a == 0:
                    cols[1] = "notanumber"
                fh.write("\t".join(cols) + "\n")
        fh.write("short\tline\n")
        # exact ties: two genera with equal weight -> the first one to appear wins
        for t in (targets[0][1], targets[1][1]):
            fh.write("\t".join(["tie_query", "1000", "0", "500", "+", t, "1", "0", "500", "500", "500", "60"]) + "\n")
    cls = load("classification_cami")
    cls.main_process(os.path.join(OUT, "hits.paf"), os.path.join(OUT, "detailed_taxonomy.tsv"),
                     os.path.join(OUT, "taxonomy_hierarchy.tsv"), os.path.join(OUT, "classified_reference.tsv"), processes=2)
    # ---- limit_candidates ----
    names = ["GCF_%09d.%d_ASM%d_genomic.fna" % (100 + i, 1 + i % 3, i) for i in range(30)] + ["weird name", "GCA_1"]
    rng.shuffle(names)
    with open(os.path.join(OUT, "selected_genomes.txt"), "w") as fh:
        fh.write("\n".join(names) + "\n\n")
    for k, fn in enumerate(("screen_a.tab", "screen_b.tab")):
        with open(os.path.join(OUT, fn), "w") as fh:
            for nm in names[k::2] + names[:5]:
                sc = rng.choice([0.99957, 0.95, 0.95, 0.9, 0.777011, 1.0])
                fh.write("%g\t%d/1000\t3\t0\t%s\tcomment\n" % (sc, rng.randrange(1, 1000), nm))
            fh.write("bad\t1/1000\t1\t0\tGCF_x\t\n\nshort\tline\n")
    os.makedirs(os.path.join(OUT, "assembly_summaries"), exist_ok=True)
    with open(os.path.join(OUT, "assembly_summaries", "assembly_summary_refseq.txt"), "w") as fh:
        fh.write("# header\n")
        for i in range(0, 30, 1):
            if i % 6 == 5:
                continue
            fh.write("\t".join(["GCF_%09d.%d" % (100 + i, 1 + i % 3), "PRJ", "SAM", "", "na", str(5000 + i), str(700 + i % 8) if i % 9 else "",
                                "Organism %d" % (i % 8), "x"]) + "\n")
    lim = load("limit_candidates")
    for tag, extra in (("plain", []), ("dedupe", ["--dedupe", "--no-download"]), ("cap", ["--max", "7"]), ("dedupe_cap", ["--dedupe", "--no-download", "--max", "4"])):
        lim.main(["--selected", os.path.join(OUT, "selected_genomes.txt"), "--output", os.path.join(OUT, "limited_%s_reference.txt" % tag),
                  "--score-file", os.path.join(OUT, "screen_a.tab"), "--score-file", os.path.join(OUT, "screen_b.tab"),
                  "--score-file", os.path.join(OUT, "missing.tab"), "--assembly-dir", os.path.join(OUT, "assembly_summaries"),
                  "--log", os.path.join(OUT, "limit_%s_reference.log" % tag)] + extra)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
