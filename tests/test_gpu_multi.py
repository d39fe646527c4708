"""N>1 on real GPUs (needs >= 2 B200 on the box; skipped otherwise): the NCCL all-reduce of
counts + all-gather of mixtures reproduce the single-process oracle bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_screen_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "dist_parity.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "parity=True" in r.stdout and "parity=False" not in r.stdout


def test_byte_ranges_and_absorb_rehearsed_on_one_gpu(tmp_path):
    """HYMET_SCREEN_GPUS's building blocks on ONE GPU: three screens over the same table each stream
    the records that start in a third of the file's bytes (hs_screen_feed_fasta_range), screen 0
    absorbs the other two (hs_screen_absorb_screen): the result equals one screen of the whole file
    and the oracle, for -w too."""
    import numpy as np

    from hymet_b200 import _lite, msh as mshfmt, screen as hs, synth
    from tests import _oracle as orc
    from tests.test_gpu_parity import build_db
    rng = np.random.default_rng(51)
    genomes = [synth.random_genome(rng, 40_000) for _ in range(30)]
    genomes += [synth.mutate(genomes[i], 0.02, rng) for i in range(5)]
    offsets, hashes, lengths = build_db(genomes, 21, 1000)
    dbp = str(tmp_path / "db.msh")
    mshfmt.write_msh(dbp, mshfmt.SketchDB(k=21, s=1000, names=[synth.gcf_name(i) for i in range(len(genomes))],
                                          comments=[""] * len(genomes), lengths=lengths, offsets=offsets, hashes=hashes))
    fasta = synth.to_fasta(synth.cut_contigs(rng, genomes[:9] + genomes[30:33], 900_000, 0.01, median=7000.0), "c")
    fap = str(tmp_path / "q.fna")
    open(fap, "wb").write(fasta)
    odb = orc.OracleDB.from_arrays(21, 1000, 42, offsets, hashes, lengths)
    db = _lite.LiteDb(dbp, 0)
    size = len(fasta)
    for wta in (False, True):
        scrs = [_lite.LiteScreen(db) for _ in range(3)]
        for i, scr in enumerate(scrs):
            scr.set_option("file_block_bytes", 65536)          # several blocks per range
            scr.feed_fasta_range(fap, size * i // 3, size * (i + 1) // 3, 2)
            scr.flush()
        parts = [scr.stats()["n_records"] for scr in scrs]
        assert sum(parts) == fasta.count(b">") and all(p > 0 for p in parts)
        scrs[0].absorb(scrs[1]); scrs[0].absorb(scrs[2])
        got = "".join(scrs[0].finish_lines(wta, 0.0, 1.0))
        want = odb.screen_text(fasta, threads=2, wta=wta)
        whole = _lite.LiteScreen(db)
        whole.feed_fasta(fap, 2)
        whole.flush()
        assert got == "".join(whole.finish_lines(wta, 0.0, 1.0)) and got.count("\n") == int((want.shared > 0).sum())
        assert scrs[0].stats()["set_size"] == want.set_size and scrs[0].stats()["n_bases"] == want.n_bases
        for scr in scrs + [whole]:
            scr.close()
    # ranges need a plain FASTA file
    import gzip
    gz = str(tmp_path / "q.fna.gz")
    gzip.open(gz, "wb").write(fasta)
    scr = _lite.LiteScreen(db)
    with pytest.raises(_lite.HsError, match="plain FASTA"):
        scr.feed_fasta_range(gz, 0, 100, 1)
    scr.close()


def test_drop_in_with_two_gpus_in_one_process(tmp_path):
    """HYMET_SCREEN_GPUS=2 bin/mash screen ... prints the same bytes as one GPU (needs >= 2 GPUs)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from hymet_b200 import msh as mshfmt, synth
    from tests.test_gpu_parity import build_db
    rng = np.random.default_rng(52)
    genomes = [synth.random_genome(rng, 40_000) for _ in range(24)]
    offsets, hashes, lengths = build_db(genomes, 21, 1000)
    dbp = str(tmp_path / "db.msh")
    mshfmt.write_msh(dbp, mshfmt.SketchDB(k=21, s=1000, names=[synth.gcf_name(i) for i in range(24)], comments=["c"] * 24,
                                          lengths=lengths, offsets=offsets, hashes=hashes))
    a, b = str(tmp_path / "a.fna"), str(tmp_path / "b.fna.gz")
    open(a, "wb").write(synth.to_fasta(synth.cut_contigs(rng, genomes[:8], 700_000, 0.01, median=6000.0), "a"))
    import gzip
    gzip.open(b, "wb").write(synth.to_fasta(synth.cut_contigs(rng, genomes[10:14], 200_000, 0.02, median=6000.0), "b"))
    exe = [sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-p", "8", "-v", "0.9"]
    for extra in ([], ["-w"]):
        one = subprocess.run(exe + extra + [dbp, a, b], capture_output=True, env=dict(os.environ, HYMET_SCREEN_GPUS="1"))
        two = subprocess.run(exe + extra + [dbp, a, b], capture_output=True, env=dict(os.environ, HYMET_SCREEN_GPUS="2"))
        assert one.returncode == 0 and two.returncode == 0, two.stderr.decode()[-2000:]
        assert one.stdout == two.stdout and one.stdout.count(b"\n") >= 5
