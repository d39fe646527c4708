#!/usr/bin/env python
"""Golden fixture from the REAL genomes the reference ships (case/truth/zymo_refs/genomes: 25 RefSeq
assemblies, strain triplets of the Zymo mock community).  Run in the build container, where
/root/reference is mounted; the outputs are committed so that the GPU box (no /root/reference) can
test against them:

  zymo25.msh          the 25 genomes sketched (k=21, s=1000, one reference per file, as `mash sketch`)
  zymo_query.fna.gz   ~0.6 Mbp of contigs cut from 12 of them (verbatim, reverse-complemented,
                      1 %-mutated), real sequence: plasmids, rRNA repeats, soft-masked and N bases
  zymo_screen.tsv     what the oracle CLI prints for `mash screen -v 0.9 zymo25.msh zymo_query.fna.gz`
  zymo_screen_w.tsv   the same with -w (near-identical strains compete for the shared hashes)

    python tests/golden/make_zymo_golden.py
"""
import glob
import gzip
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hymet_b200 import msh as mshfmt  # noqa: E402
from tests import _oracle as orc  # noqa: E402

SRC = "/root/reference/case/truth/zymo_refs/genomes"


def records(text: bytes):
    for block in text.split(b">")[1:]:
        head, _, body = block.partition(b"\n")
        yield head, body.replace(b"\n", b"").replace(b"\r", b"")


def main():
    orc.build()
    files = sorted(glob.glob(os.path.join(SRC, "*", "*.fna.gz")), key=os.path.basename)
    assert len(files) == 25, files
    names, comments, lengths, sketches, seqs = [], [], [], [], []
    for f in files:
        text = gzip.open(f, "rb").read()
        h, total = orc.sketch_text(text, 21, 1000, threads=8)
        recs = list(records(text))
        first = recs[0][0].decode()
        names.append(os.path.basename(f)[:-3])                       # mash names a sketch after its file
        comments.append("[%d seqs] %s [...]" % (len(recs), first) if len(recs) > 1 else first)
        lengths.append(total)
        sketches.append(h)
        seqs.append(recs)
    offsets = np.concatenate([[0], np.cumsum([len(h) for h in sketches])]).astype(np.uint64)
    db = mshfmt.SketchDB(k=21, s=1000, names=names, comments=comments, lengths=np.array(lengths, np.uint64),
                         offsets=offsets, hashes=np.concatenate(sketches))
    mshfmt.write_msh(os.path.join(HERE, "zymo25.msh"), db)

    rng = np.random.default_rng(25)
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    out = []
    picks = [0, 1, 3, 5, 6, 9, 12, 13, 16, 19, 22, 24]
    for n, gi in enumerate(picks):
        for c in range(5):
            head, body = seqs[gi][int(rng.integers(0, len(seqs[gi])))]
            L = int(min(len(body), rng.integers(2_000, 20_000)))
            st = int(rng.integers(0, len(body) - L + 1))
            s = bytearray(body[st:st + L])
            if c % 3 == 1:
                s = bytearray(bytes(s).translate(comp)[::-1])
            if c % 3 == 2:                                             # testdataset/mutationGCF.py model, 1 %
                for p in np.nonzero(rng.random(L) < 0.01)[0]:
                    b = chr(s[p]).upper()
                    if b in "ACGT":
                        s[p] = ord(rng.choice([x for x in "ACGT" if x != b]))
            out.append(b">contig_%d_%d from %s\n" % (n, c, names[gi].encode()))
            out += [bytes(s[i:i + 70]) + b"\n" for i in range(0, len(s), 70)]
    query = b"".join(out)
    qpath = os.path.join(HERE, "zymo_query.fna.gz")
    with gzip.GzipFile(qpath, "wb", mtime=0) as fh:
        fh.write(query)
    for extra, name in (([], "zymo_screen.tsv"), (["-w"], "zymo_screen_w.tsv")):
        r = subprocess.run([orc.BIN, "screen", "-p", "4", "-v", "0.9"] + extra + [os.path.join(HERE, "zymo25.msh"), qpath],
                           capture_output=True, check=True)
        open(os.path.join(HERE, name), "wb").write(r.stdout)
        print(name, r.stdout.count(b"\n"), "lines")
    print("query bases", sum(len(l) - 1 for l in out if not l.startswith(b">")), "gz bytes", os.path.getsize(qpath),
          "msh bytes", os.path.getsize(os.path.join(HERE, "zymo25.msh")))


if __name__ == "__main__":
    main()
