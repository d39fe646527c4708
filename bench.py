#!/usr/bin/env python
"""bench.py -- `mash screen` query throughput (Mbp/s) on B200, per the driver contract.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  (one rank per GPU)

Workload (BASELINE.json configs[1]): synthetic 1 Gbp CAMI-shaped contig set (contigs cut
from 500 x 2 Mb genomes at 1 % substitutions, half reverse-complemented, Zymo-fitted
length distribution) vs a 50 000-sketch database (k=21, s=1000; 500 sketched from the
genomes on the GPU + 49 500 decoy sketches).  One step = one complete screen of the
rank's contig shard: reset, stream (k-mer hash + probe + mixture), mixture bottom-s,
[N>1: one NCCL all-reduce of counts + all-gather of mixtures], per-sketch reduction,
identity/p-value, results back on the host.

  value  : device-resident (packed query already in HBM), whole job, CUDA-event timed,
           max over ranks.
  e2e    : same screen through the public API from FASTA TEXT in pinned host memory:
           host 2-bit packing + H2D + kernels + D2H inside the timed region.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mash-screen query throughput"
UNIT = "Mbp/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mbp", type=int, default=1000, help="query Mbp per GPU")
    ap.add_argument("--sketches", type=int, default=50_000)
    ap.add_argument("--real", type=int, default=500, help="sketches made from real (synthetic) genomes")
    ap.add_argument("--cpu-mbp", type=int, default=48, help="bounded sample for the CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-filter", action="store_true", help="probe the table for every k-mer (mash semantics, no range pre-filter)")
    ap.add_argument("--wta", action="store_true")
    ap.add_argument("--tiny", type=int, default=0,
                    help="this many decoy sketches come from tiny genomes (1.5-20 k k-mers): their hashes spread "
                         "over the whole range and defeat the range pre-filter (viral/plasmid-heavy databases)")
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--s", type=int, default=1000)
    ap.add_argument("--clusters", type=int, default=0,
                    help="config 4: make this many of the real genomes 1-5 %% diverged copies of the others")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_from_fasta(fasta_np, want_bytes: int) -> bytes:
    """Prefix of the FASTA text cut at a record boundary (bounded CPU-baseline sample)."""
    n = len(fasta_np)
    if want_bytes >= n:
        return fasta_np.tobytes()
    tail = fasta_np[want_bytes:min(n, want_bytes + (8 << 20))].tobytes()
    j = tail.find(b"\n>")
    end = n if j < 0 else want_bytes + j + 1
    return fasta_np[:end].tobytes()


def run_reference(args):
    """CPU arm: the oracle restatement of `mash screen` (the real binary is a third-party
    dependency that is not under /root/reference nor installable offline -- DESIGN.md),
    all host threads, on a bounded sample of the same workload shape."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from hymet_b200 import synth
    from tests import _oracle as orc

    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    k, s = 21, 1000
    rng = np.random.default_rng(2)
    n_real = 24
    genomes = [synth.random_genome(rng, 2_000_000) for _ in range(n_real)]
    real = np.stack([orc.sketch_text(synth.to_fasta([g], "g", width=0), k, s, threads=threads)[0] for g in genomes])
    decoy, dlen = synth.decoy_sketches(rng, args.sketches - n_real, s)
    hashes = np.concatenate([real.reshape(-1), decoy.reshape(-1)])
    lengths = np.concatenate([np.full(n_real, 2_000_000, np.uint64), dlen])
    offsets = np.arange(args.sketches + 1, dtype=np.uint64) * np.uint64(s)
    t0 = time.perf_counter()
    odb = orc.OracleDB.from_arrays(k, s, 42, offsets, hashes, lengths)
    t_table = time.perf_counter() - t0
    fasta = synth.to_fasta(synth.cut_contigs(rng, genomes, args.cpu_mbp * 1_000_000, 0.01), "c", width=80)
    times, bases = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = odb.screen_text(fasta, threads=threads, wta=args.wta)
        dt = time.perf_counter() - t0
        bases = r.n_bases
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = args.steps * bases / total / 1e6
    sample = "%d Mbp of contigs (%d bases) vs the full %d-sketch table, per step" % (args.cpu_mbp, bases, args.sketches)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        # same workload name and keys as the CUDA arm's line; the sample this arm timed is in cpu_baseline.sample
        "config": {"workload": "c2: synthetic %d Mbp CAMI-shaped contig set per GPU vs %d-sketch db (k=%d, s=%d)"
                               % (args.mbp, args.sketches, k, s), "k": k, "s": s,
                   "query_mbp_per_gpu": args.mbp, "sketches": args.sketches, "mutation_rate": 0.01,
                   "winner_take_all": bool(args.wta),
                   "cpu_sample_mbp_per_step": args.cpu_mbp,
                   "timing": "host wall clock around the oracle's screen call, all host threads"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "table_build_s": t_table,
                         "note": "oracle/mash_screen_oracle.c (mash-semantics restatement); real mash is absent from this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from hymet_b200 import dist as hd
    from hymet_b200 import screen as hs
    from hymet_b200 import workload

    rank, world, local = hd.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    t_setup = time.perf_counter()
    want_host = not args.no_e2e
    wl = workload.make_c2(local, mbp=args.mbp, n_sketches=args.sketches, n_real=args.real, shard=rank,
                          with_fasta=want_host, with_host_packed=want_host, k=args.k, s=args.s,
                          cluster_copies=args.clusters, tiny=args.tiny)
    db = hs.Database.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths, device=local)
    stream = torch.cuda.Stream(device=dev)   # explicit stream: the library launches on it, the events time it
    torch.cuda.set_stream(stream)
    scr = hd.DistributedScreen(db, local, stream_ptr=stream.cuda_stream, probe_filter=not args.no_filter)
    t_setup = time.perf_counter() - t_setup
    n_cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    host_threads = max(1, n_cpus // world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        scr.reset()
        scr.feed_packed_device(wl.d_seq.data_ptr(), wl.d_inv.data_ptr(), wl.n_positions)
        return scr.finish(args.wta)

    text_wall = {"reset": 0.0, "feed": 0.0, "finish": 0.0, "n": 0}

    def step_text():
        t0 = time.perf_counter()
        scr.reset()
        t1 = time.perf_counter()
        scr.feed_text_ptr(wl.fasta.data_ptr(), wl.fasta.numel(), host_threads)
        t2 = time.perf_counter()
        r = scr.finish(args.wta)
        t3 = time.perf_counter()
        text_wall.setdefault("steps", []).append([round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2),
                                                  round(1e3 * (t3 - t2), 2)])
        return r

    def step_packed_host():
        scr.reset()
        scr.feed_packed_ptr(wl.h_seq.data_ptr(), wl.h_inv.data_ptr(), wl.n_positions)
        return scr.finish(args.wta)

    def timed(fn, steps, warmup, sample_clocks=False):
        res = None
        for _ in range(warmup):
            res = fn()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats = []
        if sample_clocks:
            torch.cuda.profiler.start()   # `ncu --profile-from-start off` sees only the timed region
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
            stats.append(res.stats)
        e1.record(stream)
        barrier()
        if sample_clocks:
            torch.cuda.profiler.stop()
        wall = time.perf_counter() - t0
        n_timed = len(sampler.rows) if sampler else 0
        if sampler and n_timed < 25:
            # the timed region is only tens of ms: keep the same load running (untimed) until the
            # sampler has seen it for ~0.7 s, so the median clock is of the kernel, not of idle
            t_end = time.perf_counter() + 0.7
            while time.perf_counter() < t_end:
                fn()
            torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if clocks is not None:
            clocks["samples_in_timed_region"] = n_timed
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, stats, res, clocks

    total_bases = wl.n_bases
    if world > 1:
        t = torch.tensor([wl.n_bases], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        total_bases = int(t[0])

    ms, wall, stats, res, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    value = args.steps * total_bases / (ms * 1e-3) / 1e6
    st = stats[-1]
    launches = sum(s["n_launches"] for s in stats)

    # ---- roofline of the dominant kernel (k_stream), from its own CUDA events -------
    peak, peak_src = measured_peak_gbs()
    ms_stream = sum(s["ms_stream"] for s in stats) / len(stats)
    ms_reduce = sum(s["ms_reduce"] for s in stats) / len(stats)
    ms_reset = sum(s["ms_reset"] for s in stats) / len(stats)
    alg_bytes = st["n_positions"] * 3 / 8 + 128 * st["n_bucket_reads"] + 64 * st["n_hits"]
    sem_bytes = st["n_positions"] * 3 / 8 + 128 * st["n_valid_kmers"] + 64 * st["n_hits"]
    achieved = alg_bytes / (ms_stream * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = prof.get("k_stream", {})
        if ent.get("mbp") == args.mbp and ent.get("sketches") == args.sketches and bool(ent.get("filter", True)) == (not args.no_filter):
            traffic = ent.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "k_stream<%d>" % args.k, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "ms_per_launch": ms_stream, "share_of_step": ms_stream / (ms / args.steps),
                "algorithmic_bytes_per_launch": alg_bytes,
                "semantic_bytes_per_launch": sem_bytes,
                "semantic_frac": sem_bytes / (ms_stream * 1e-3) / 1e9 / peak,
                "kmers_per_s": st["n_valid_kmers"] / (ms_stream * 1e-3),
                "note": "k_stream is integer-issue bound (MurmurHash3 per k-mer); the exact range pre-filter removes "
                        "almost all probe traffic, so the HBM fraction is low by design -- see DESIGN.md 'Roofline'"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "c2: synthetic %d Mbp CAMI-shaped contig set per GPU vs %d-sketch db (k=%d, s=%d)"
                               % (args.mbp, args.sketches, args.k, args.s), "k": args.k, "s": args.s,
                   "query_mbp_per_gpu": args.mbp, "contigs_per_gpu": wl.n_contigs, "sketches": args.sketches,
                   "real_genome_sketches": wl.n_real, "mutation_rate": 0.01, "probe_filter": not args.no_filter,
                   "winner_take_all": bool(args.wta), "parallelism": "query sharded x%d, table replicated" % world,
                   "count_exchange": scr.last_exchange,
                   "l2": "inputs larger than L2 (%.0f MB packed query + %.0f MB table per GPU)"
                         % (wl.n_positions * 0.375 / 1e6, db.info.device_bytes / 1e6),
                   "timing": "CUDA events on the launching stream, max over ranks"},
        "roofline": roofline,
        "gpu_launches": launches,
        "clocks": clocks,
        "step_breakdown_ms": {"stream_kernel": ms_stream, "mixture_and_reduce": ms_reduce, "reset": ms_reset,
                              "host_gaps_and_result_copy": ms / args.steps - ms_stream - ms_reduce - ms_reset,
                              "step_total": ms / args.steps, "wall_per_step": 1e3 * wall / args.steps},
        "reduce": {"path": "dense O(stored hashes)" if st["reduce_path"] else "sparse O(present hashes)",
                   "present_hashes": st["n_touched"], "refs_with_hits": st["n_hit_refs"], "pairs_walked": st["n_pairs"]},
        "counters": {k_: st[k_] for k_ in ("n_positions", "n_valid_kmers", "n_probes", "n_bucket_reads", "n_hits",
                                          "n_mix_inserts", "n_mix_passes", "set_size")},
        "db": {"distinct_hashes": int(db.n_distinct), "table_mb": db.info.device_bytes / 1e6,
               "tiny_genome_sketches": args.tiny, "bloom_mb": db.info.bloom_bytes / 1e6,
               "range_filter_pass_fraction": db.info.max_key / 2.0 ** 64,
               "dense_range_fraction": db.info.dense_max / 2.0 ** 64, "keys_in_bloom_tier": int(db.info.bloom_keys),
               "table_build_s": db.info.t_build_s},
        "setup_s": t_setup,
    }

    # ---- K2 alone: random probes against the HBM-resident table (the north-star probe roofline)
    if rank == 0:
        n_probe = 1 << 26
        hq = torch.randint(-(1 << 62), 1 << 62, (n_probe,), dtype=torch.int64, device=dev)
        best = None
        for _ in range(4):
            hits, reads, pms = db.probe_device(hq.data_ptr(), n_probe)
            best = pms if best is None else min(best, pms)
        # what this GPU delivers for raw random 32-byte sector reads (no hashing), same footprint
        gbuf = torch.empty(int(db.info.n_buckets) * 16, dtype=torch.int64, device=dev)
        g_ms = min(hs.gather_bench(gbuf.data_ptr(), gbuf.numel() * 8, n_probe) for _ in range(3))
        del gbuf
        p_traffic, p_src = None, None
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("k_probe_r02", {})
            if ent.get("probes") == n_probe and ent.get("sketches") == args.sketches:
                p_traffic, p_src = ent.get("dram_bytes_per_launch"), "static: " + ent.get("source", "profiles/")
        except Exception:
            pass
        gbs = (128.0 * reads + 8.0 * n_probe) / (best * 1e-3) / 1e9
        line["probe_kernel"] = {"kernel": "k_probe (warp-cooperative, 8 lanes per 128-byte bucket)", "probes": n_probe,
                                "bucket_reads": int(reads), "ms": best,
                                "probes_per_s": n_probe / (best * 1e-3), "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                "frac": gbs / peak,
                                "bytes": "128 B bucket line per read (10 keys + their ids + overflow flag: everything a probe "
                                         "needs, hit or miss) + 8 B hash read per probe",
                                "traffic": p_traffic, "traffic_source": p_src,
                                "dram_frac_with_measured_traffic": (p_traffic / (best * 1e-3) / 1e9 / peak) if p_traffic else None,
                                "table_bytes_per_key": db.info.n_buckets * 128.0 / max(1, db.info.n_entries),
                                "round1_layout": "32 B buckets of 4 keys + ids in a second array: 4.0e10 probes/s, 0.25 of peak in "
                                                 "useful bytes, 0.78 in moved bytes (B200 fills 128 B per 32 B read), 36 B per key",
                                "random_sector_reads_per_s_of_this_gpu": n_probe / (g_ms * 1e-3),
                                "frac_of_random_sector_rate": (reads / (best * 1e-3)) / (n_probe / (g_ms * 1e-3))}
        del hq

    # ---- e2e: FASTA text in pinned host memory -> TSV columns on the host --------------
    if not args.no_e2e:
        e_steps = max(1, min(args.steps, 3))
        # three ways to get FASTA text into HBM (option "ingest"): host AVX2 packer threads,
        # device parser fed by DMA of the raw text, or both competing for chunks (the default)
        modes = {}
        for mode, name in ((0, "host_packer"), (1, "device_parser"), (2, "hybrid")):
            scr.set_option("ingest", mode)
            mms, _, mstats, mres, _ = timed(step_text, e_steps, 2)
            modes[name] = {"value": e_steps * total_bases / (mms * 1e-3) / 1e6, "ms_per_step": mms / e_steps,
                           "h2d_bytes_per_step": int(mstats[-1]["h2d_bytes"]),
                           "ok": bool(mres.shared.tolist() == res.shared.tolist() and mres.set_size == res.set_size)}
        scr.set_option("ingest", 2)
        ems, ewall, estats, eres, _ = timed(step_text, e_steps, 1)
        est = estats[-1]
        e_val = e_steps * total_bases / (ems * 1e-3) / 1e6
        line["e2e"] = {"value": e_val, "unit": UNIT, "h2d_bytes_per_step": int(est["h2d_bytes"]),
                       "d2h_bytes_per_step": int(est["d2h_bytes"]), "ms_per_step": ems / e_steps,
                       "input": "FASTA text (%d B per GPU, 80-column lines) in pinned host memory" % wl.fasta.numel(),
                       "host_threads_per_gpu": host_threads, "steps": e_steps, "ingest": "hybrid (default)",
                       "ingest_modes": modes,
                       "host_wall_ms_reset_feed_finish": text_wall["steps"][-e_steps:],
                       "stream_kernel_ms_per_step": est["ms_stream"], "launches_per_step": est["n_launches"],
                       "includes": "host FASTA parse + 2-bit pack, H2D, all kernels, D2H of the result columns"}
        pms, pwall, pstats, pres, _ = timed(step_packed_host, e_steps, 1)
        line["e2e_packed"] = {"value": e_steps * total_bases / (pms * 1e-3) / 1e6, "unit": UNIT,
                              "h2d_bytes_per_step": int(pstats[-1]["h2d_bytes"]),
                              "d2h_bytes_per_step": int(pstats[-1]["d2h_bytes"]), "ms_per_step": pms / e_steps,
                              "input": "pre-packed 2-bit + mask words in pinned host memory"}
        # SURVEY 8d's end-to-end: the FASTA is a FILE in the page cache, read by hs_screen_feed_fasta (reader
        # threads pread() record-aligned blocks into a pinned ring, the device parses them)
        import tempfile
        fdir = tempfile.mkdtemp(prefix="hs_bench_%d_" % rank)
        fpath = os.path.join(fdir, "contigs.fna")
        wl.fasta.numpy().tofile(fpath)
        readers = max(1, min(8, host_threads))
        scr.set_option("file_readers", readers)
        scr.set_option("file_block_bytes", 16 << 20)

        def step_file():
            scr.reset()
            scr.feed_fasta(fpath, host_threads)
            return scr.finish(args.wta)

        fms, fwall, fstats, fres, _ = timed(step_file, e_steps, 2)
        line["e2e_file"] = {"value": e_steps * total_bases / (fms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": fms / e_steps,
                            "h2d_bytes_per_step": int(fstats[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(fstats[-1]["d2h_bytes"]),
                            "input": "the same FASTA as a file in the page cache (%d B per GPU)" % wl.fasta.numel(),
                            "reader_threads_per_gpu": readers,
                            "ok": bool(fres.shared.tolist() == res.shared.tolist() and fres.set_size == res.set_size)}
        try:
            os.remove(fpath)
            os.rmdir(fdir)
        except OSError:
            pass
        # the three entry points must agree exactly
        same = (eres.shared.tolist() == res.shared.tolist() and eres.median.tolist() == res.median.tolist()
                and pres.shared.tolist() == res.shared.tolist() and eres.set_size == res.set_size)
        line["e2e"]["matches_device_resident"] = bool(same)
    else:
        line["e2e"] = None

    # ---- CPU baseline on a bounded sample (rank 0, N=1) + parity of that sample ----------
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_e2e:
        from tests import _oracle as orc
        threads = n_cpus
        sample = cpu_sample_from_fasta(wl.fasta.numpy(), int(args.cpu_mbp * 1_000_000 * 82 / 80))
        t0 = time.perf_counter()
        odb = orc.OracleDB.from_arrays(wl.k, wl.s, 42, wl.offsets, wl.hashes, wl.lengths)
        t_table = time.perf_counter() - t0
        best, r = None, None
        for _ in range(2):
            t0 = time.perf_counter()
            r = odb.screen_text(sample, threads=threads, wta=args.wta)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        scr.reset()
        scr.feed_text(sample, host_threads)
        g = scr.finish(args.wta)
        ok = (g.shared.tolist() == r.shared.tolist() and g.median.tolist() == r.median.tolist()
              and g.set_size == r.set_size
              and bool(np.all(np.abs(g.identity - r.identity) <= 1e-12 * np.abs(r.identity)))
              and bool(np.all(np.abs(g.pvalue - r.pvalue) <= 1e-12 * np.abs(r.pvalue))))
        line["cpu_baseline"] = {"value": r.n_bases / best / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "first %d bases of the same contig set vs the same %d-sketch table, best of 2"
                                          % (r.n_bases, args.sketches),
                                "table_build_s": t_table, "parity_on_sample": "bit-exact" if ok else "MISMATCH",
                                "sample_shared_hashes": int(r.shared.sum())}
    if rank == 0:
        emit(line)
    scr.scr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_OUT = None


def emit(line) -> None:
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    args = parse_args()
    # libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
