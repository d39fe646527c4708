#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): sharded multi-GPU screen == single-process oracle.
Exit status 0 on parity.  Used by tests/test_gpu_multi.py and by hand:
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hymet_b200 import dist as hd, screen as hs, synth  # noqa: E402
from tests import _oracle as orc  # noqa: E402


def main():
    rank, world, local = hd.init_from_env("nccl")
    torch.cuda.set_device(local)
    rng = np.random.default_rng(31)
    k, s = 21, 1000
    genomes = [synth.random_genome(rng, 60_000) for _ in range(40)]
    genomes += [synth.mutate(genomes[i], 0.02, rng) for i in range(6)]        # near-duplicates for -w
    sk = [orc.sketch_text(synth.to_fasta([g], "g"), k, s)[0] for g in genomes]
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in sk])]).astype(np.uint64)
    hashes = np.concatenate(sk)
    lengths = np.array([len(g) - (i % 5) for i, g in enumerate(genomes)], np.uint64)
    fasta = synth.to_fasta(synth.cut_contigs(rng, genomes[:12], 3_000_000, 0.01, median=6000.0), "c")
    db = hs.Database.from_arrays(k, s, 42, offsets, hashes, lengths, device=local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ok = True
    runs = [(False, "dense", None), (True, "dense", None), (False, "sparse", None), (True, "sparse", None)]
    auto = hd.DistributedScreen(db, local, exchange="auto", stream_ptr=stream.cuda_stream)
    # the same handle three times: the record is sized from the previous screen (first: too small -> dense
    # fallback, then sparse)
    runs += [(False, "auto", auto), (True, "auto", auto), (False, "auto", auto)]
    # the sync-free flush's fallback: ONE rank's device-side selection is made to report "not settled";
    # every rank must notice (the verdict travels in the gathered records) and redo the mixture exchange
    if rank == world - 1:
        os.environ["HYMET_SCREEN_FORCE_UNSETTLED"] = "1"
    forced = hd.DistributedScreen(db, local, exchange="sparse", stream_ptr=stream.cuda_stream)
    os.environ.pop("HYMET_SCREEN_FORCE_UNSETTLED", None)
    runs += [(True, "sparse+unsettled", forced), (False, "sparse+settled", forced)]
    for wta, mode, handle in runs:
        scr = handle or hd.DistributedScreen(db, local, exchange=mode, stream_ptr=stream.cuda_stream)
        if handle:
            scr.reset()
        b, e = hd.record_aligned_range(fasta, rank, world)
        scr.feed_text(fasta[b:e], 2)
        res = scr.finish(wta)
        if rank == 0:
            want = orc.OracleDB.from_arrays(k, s, 42, offsets, hashes, lengths).screen_text(fasta, threads=4, wta=wta)
            same = (res.shared.tolist() == want.shared.tolist() and res.median.tolist() == want.median.tolist()
                    and res.set_size == want.set_size and scr.mixture().tolist() == want.mixture.tolist()
                    and bool(np.all(np.abs(res.identity - want.identity) <= 1e-12 * np.abs(want.identity)))
                    and bool(np.all(np.abs(res.pvalue - want.pvalue) <= 1e-12 * np.abs(want.pvalue))))
            print("world=%d wta=%d exchange=%s(%s) shared_total=%d parity=%s" % (world, wta, mode, scr.last_exchange,
                                                                                  int(res.shared.sum()), same), flush=True)
            ok = ok and same
        if mode == "sparse+unsettled" and scr.n_unsettled != 1:
            print("rank %d: the forced unsettled selection was not noticed (n_unsettled=%d)" % (rank, scr.n_unsettled), flush=True)
            ok = False
        if mode == "sparse+settled" and scr.n_unsettled != 1:
            print("rank %d: a settled screen fell back (n_unsettled=%d)" % (rank, scr.n_unsettled), flush=True)
            ok = False
        if not handle:
            scr.scr.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
