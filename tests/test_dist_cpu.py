"""Multi-GPU host logic on CPU: query sharding + the two exchanges, world_size 2, gloo.
The all-reduce and all-gather are exercised for real (gloo); the per-rank screens are
played by the oracle so that rank-sharded + exchanged == single-process, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hymet_b200 import dist as hd
from hymet_b200 import synth
from tests import _oracle as orc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _case():
    rng = np.random.default_rng(21)
    genomes = [synth.random_genome(rng, 20_000) for _ in range(12)]
    sk = [orc.sketch_text(synth.to_fasta([g], "g"), 21, 200)[0] for g in genomes]
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in sk])]).astype(np.uint64)
    fasta = synth.to_fasta(synth.cut_contigs(rng, genomes[:5], 150_000, 0.01, median=3000.0), "c")
    return offsets, np.concatenate(sk), np.full(12, 20_000, np.uint64), fasta


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = hd.init_from_env("gloo")
    offsets, hashes, lengths, fasta = _case()
    db = orc.OracleDB.from_arrays(21, 200, 42, offsets, hashes, lengths)
    b, e = hd.record_aligned_range(fasta, r, w)
    local = db.screen_text(fasta[b:e]) if e > b else None
    counts = torch.from_numpy((local.counts_per_entry if local is not None else np.zeros(len(hashes), np.uint32))
                              .view(np.int32).copy())
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    parts = hd.all_gather_mixture(local.mixture if local is not None else np.zeros(0, np.uint64), 200)
    merged = hd.merge_bottom_s(parts, 200)
    q.put((r, (b, e), counts.numpy().view(np.uint32).copy(), merged))
    dist.barrier()
    dist.destroy_process_group()


def _absorb_cpu(counts, rows, cap, skip):
    """CPU stand-in for hs_screen_counts_absorb (k_counts_absorb): add every other rank's pairs, or
    nothing at all when any record is incomplete.  Returns (overflow, largest pair count)."""
    n_pairs = (rows[:, 0] & 0xFFFFFFFF).astype(np.int64)
    most = int(n_pairs.max())
    if most > cap:
        return True, most
    for rr in range(rows.shape[0]):
        if rr != skip:
            i2, c2 = hd.unpack_pairs(rows[rr, 1:1 + int(n_pairs[rr])])
            np.add.at(counts, i2, c2)
    return False, most


def _record_worker(rank, world, port, q):
    """The exchange of hymet_b200.dist.DistributedScreen with CPU stand-ins for the device kernels
    (pair compaction, absorb, mixture merge): same record layouts, same helpers, same capacity policy,
    same fallback decision taken from the gathered records alone."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = hd.init_from_env("gloo")
    offsets, hashes, lengths, fasta = _case()
    s, E = 200, len(hashes)
    db = orc.OracleDB.from_arrays(21, s, 42, offsets, hashes, lengths)
    b, e = hd.record_aligned_range(fasta, r, w)
    local = db.screen_text(fasta[b:e])
    cap, log = 16, []                      # far too small: the first screen must fall back to the dense sum
    for step in range(3):
        counts = local.counts_per_entry.astype(np.uint32).copy()
        ids = np.nonzero(counts)[0]
        rec = np.zeros(1 + cap, np.int64)
        rec[0] = len(ids)                                               # pair count, then <= cap pairs
        rec[1:1 + min(cap, len(ids))] = hd.pack_pairs(ids[:cap], counts[ids[:cap]])
        out = torch.empty(w * len(rec), dtype=torch.int64)
        work = dist.all_gather_into_tensor(out, torch.from_numpy(rec), async_op=True)
        mrec = hd.mixture_record(local.mixture, s)
        mout = torch.empty(w * len(mrec), dtype=torch.int64)
        dist.all_gather_into_tensor(mout, torch.from_numpy(mrec))
        merged = hd.merge_bottom_s(hd.parse_mixture_rows(mout.numpy().reshape(w, 1 + s)), s)
        work.wait()
        overflow, most = _absorb_cpu(counts, out.numpy().reshape(w, 1 + cap), cap, r)
        new_cap = hd.next_cap(most, cap, E)
        mode = "sparse"
        if overflow:
            t = torch.from_numpy(counts.view(np.int32))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            mode = "dense"
        log.append((mode, cap, counts.copy(), merged))
        cap = new_cap
    q.put((r, log))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sparse_record_exchange_protocol_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_record_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    offsets, hashes, lengths, fasta = _case()
    whole = orc.OracleDB.from_arrays(21, 200, 42, offsets, hashes, lengths).screen_text(fasta)
    for r in range(world):
        modes = [m for m, _, _, _ in got[r]]
        assert modes == ["dense", "sparse", "sparse"]                 # overflow -> fallback, then the adapted record
        assert got[r][1][1] >= 4096 and got[r][2][1] == got[r][1][1]  # capacity settles
        for mode, cap, counts, merged in got[r]:
            assert np.array_equal(counts, whole.counts_per_entry), mode
            assert np.array_equal(merged, whole.mixture)
    assert [c for _, c, _, _ in got[0]] == [c for _, c, _, _ in got[1]]   # both ranks take the same decisions


def test_record_helpers():
    assert hd.next_cap(0, 4096, 10 ** 6) == 4096
    assert hd.next_cap(276_000, 1 << 20, 5 * 10 ** 7) == 393_216          # C2: a quarter million pairs -> 3 MB records
    assert hd.next_cap(276_000, 393_216, 5 * 10 ** 7) == 393_216          # stays
    assert hd.next_cap(300_000, 393_216, 5 * 10 ** 7) == 393_216          # ... also when the count moves a little
    assert hd.next_cap(600_000, 393_216, 5 * 10 ** 7) == 786_432          # overflow: grows
    assert hd.next_cap(1_467_217, 1 << 22, 3 * 10 ** 8) == 1_835_008      # C3 at two GPUs: 14 MB instead of 32
    assert hd.next_cap(10 ** 6, 4096, 5000) == 5000                        # never more than one pair per entry
    ids = np.array([0, 7, 2 ** 31 + 5], np.int64); cnt = np.array([1, 2 ** 32 - 1, 9], np.uint32)
    i2, c2 = hd.unpack_pairs(hd.pack_pairs(ids, cnt))
    assert i2.tolist() == ids.tolist() and c2.tolist() == cnt.tolist()
    mix = np.array([3, 2 ** 63 + 1], np.uint64)
    m = hd.parse_mixture_rows(np.stack([hd.mixture_record(mix, 4), hd.mixture_record(np.zeros(0, np.uint64), 4)]))
    assert m[0].tolist() == mix.tolist() and len(m[1]) == 0


@pytest.mark.timeout(120)
def test_sharded_counts_and_mixture_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    offsets, hashes, lengths, fasta = _case()
    whole = orc.OracleDB.from_arrays(21, 200, 42, offsets, hashes, lengths).screen_text(fasta)
    ranges = sorted(g[1] for g in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(fasta) and ranges[0][1] == ranges[1][0]
    for _, _, counts, merged in got:
        assert np.array_equal(counts, whole.counts_per_entry)      # all-reduce(sum) == unsharded counts
        assert np.array_equal(merged, whole.mixture)               # merged bottom-s == unsharded bottom-s
    assert orc.lib().orc_set_size(got[0][3].ctypes.data_as(orc.C.POINTER(orc.C.c_uint64)), len(got[0][3]), 1) == whole.set_size


def test_record_aligned_ranges_cover_without_splitting_records():
    rng = np.random.default_rng(5)
    fasta = synth.to_fasta([synth.random_genome(rng, int(n)) for n in rng.integers(10, 4000, size=37)], "r")
    for world in (1, 2, 3, 8, 64):
        rs = [hd.record_aligned_range(fasta, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == len(fasta)
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        assert all(b == e or fasta[b:b + 1] == b">" for b, e in rs)
    fq = b"@r1\nACGT\n+\n!!!!\n@r2\nGGCC\n+\n@@@@\n"
    assert hd.record_aligned_range(fq, 0, 2) == (0, len(fq)) and hd.record_aligned_range(fq, 1, 2) == (len(fq), len(fq))


def test_merge_bottom_s():
    a = np.array([1, 5, 9], np.uint64); b = np.array([2, 5, 7, 11], np.uint64)
    assert hd.merge_bottom_s([a, b], 4).tolist() == [1, 2, 5, 7]
    assert hd.merge_bottom_s([a, np.zeros(0, np.uint64)], 10).tolist() == [1, 5, 9]
