"""Gbp-scale synthetic workloads built on the GPU (torch is plumbing here: device
memory + random numbers).  Same generators as hymet_b200.synth (mutationGCF.py model,
Zymo-fitted contig lengths), vectorised so a 1 Gbp contig set takes seconds.

Shapes follow BASELINE.json `configs` / SURVEY.md 8d:
  c2: 50 000 sketches (k=21, s=1000) = `n_real` genomes of 2 Mb sketched from their
      sequence with the GPU sketcher + decoy sketches (bottom-s of uniform hashes),
      query = `mbp` Mbp of contigs cut from the real genomes at 1 % substitutions,
      half of them reverse-complemented, one record per contig.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import screen as hs
from . import synth


@dataclass
class Workload:
    name: str
    k: int
    s: int
    offsets: np.ndarray          # sketch db, flat
    hashes: np.ndarray
    lengths: np.ndarray
    n_real: int
    d_seq: torch.Tensor          # packed query, device (int64 words)
    d_inv: torch.Tensor          # int32 words
    n_positions: int             # packed positions (bases + one separator per contig)
    n_bases: int                 # bases inside records: the numerator of Mbp/s
    n_contigs: int
    fasta: Optional[torch.Tensor] = None   # pinned uint8 host tensor with the same contigs as FASTA text
    h_seq: Optional[torch.Tensor] = None   # pinned packed copies (e2e from packed host buffers)
    h_inv: Optional[torch.Tensor] = None
    fasta_sample: Optional[bytes] = None   # FASTA text of the first contigs (bounded CPU-oracle sample)


def _contig_table(rng: np.random.Generator, n_genomes: int, genome_len: int, total: int):
    lens = np.array(synth.contig_lengths(rng, total), dtype=np.int64)
    lens = np.minimum(lens, genome_len)
    g = rng.integers(0, n_genomes, size=len(lens))
    st = (rng.random(len(lens)) * (genome_len - lens + 1)).astype(np.int64)
    rc = rng.random(len(lens)) < 0.5
    return lens, g, st, rc


def _fill_query_codes(genomes: torch.Tensor, genome_len: int, lens, g, st, rc, rate: float, gen: torch.Generator,
                      block: int = 1 << 26) -> torch.Tensor:
    """codes[P] uint8: 4 at the first position of every contig (record separator), then its bases."""
    dev = genomes.device
    span = lens + 1
    qoff = np.concatenate([[0], np.cumsum(span)])
    P = int(qoff[-1])
    out = torch.empty(P, dtype=torch.uint8, device=dev)
    c0 = 0
    while c0 < len(lens):
        c1 = c0
        while c1 < len(lens) and qoff[c1 + 1] - qoff[c0] <= block:
            c1 += 1
        c1 = max(c1, c0 + 1)
        sp = torch.from_numpy(span[c0:c1]).to(dev)
        cid = torch.repeat_interleave(torch.arange(c1 - c0, device=dev), sp)
        base = torch.from_numpy(qoff[c0:c1] - qoff[c0]).to(dev)
        off = torch.arange(int(qoff[c1] - qoff[c0]), device=dev) - base[cid]   # 0 = separator
        L = torch.from_numpy(lens[c0:c1]).to(dev)[cid]
        start = torch.from_numpy(g[c0:c1] * genome_len + st[c0:c1]).to(dev)[cid]
        is_rc = torch.from_numpy(rc[c0:c1]).to(dev)[cid]
        pos = off - 1
        src = torch.where(is_rc, start + (L - 1 - pos), start + pos).clamp_(min=0)
        code = genomes[src]
        code = torch.where(is_rc, 3 - code, code)
        # mutationGCF.py:4-18: with probability `rate` replace by a uniformly chosen different base
        hit = torch.rand(code.shape, device=dev, generator=gen) < rate
        shift = torch.randint(1, 4, code.shape, device=dev, generator=gen, dtype=torch.uint8)
        code = torch.where(hit, (code + shift) % 4, code)
        code = torch.where(off == 0, torch.full_like(code, 4), code)
        out[int(qoff[c0]):int(qoff[c1])] = code
        del cid, off, L, start, is_rc, pos, src, code, hit, shift
        c0 = c1
    return out


def _fasta_from_codes(codes: torch.Tensor, lens: np.ndarray, width: int, block: int = 1 << 26) -> torch.Tensor:
    """FASTA text (uint8, pinned host tensor): '>c%09d\\n' + bases wrapped at `width` + '\\n' per contig."""
    dev = codes.device
    H = 12  # len(">c%09d\n")
    nlines = (lens + width - 1) // width
    tlen = H + lens + nlines
    toff = np.concatenate([[0], np.cumsum(tlen)])
    qoff = np.concatenate([[0], np.cumsum(lens + 1)])
    T = int(toff[-1])
    host = torch.empty(T, dtype=torch.uint8, pin_memory=True)
    hdr = np.frombuffer(b"".join(b">c%09d\n" % i for i in range(len(lens))), dtype=np.uint8).reshape(len(lens), H)
    hdr_d = torch.from_numpy(hdr.copy()).to(dev)
    ascii_lut = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device=dev)
    c0 = 0
    while c0 < len(lens):
        c1 = c0
        while c1 < len(lens) and toff[c1 + 1] - toff[c0] <= block:
            c1 += 1
        c1 = max(c1, c0 + 1)
        tl = torch.from_numpy(tlen[c0:c1]).to(dev)
        cid = torch.repeat_interleave(torch.arange(c1 - c0, device=dev), tl)
        off = torch.arange(int(toff[c1] - toff[c0]), device=dev) - torch.from_numpy(toff[c0:c1] - toff[c0]).to(dev)[cid]
        body = off - H                       # >= 0 inside the wrapped sequence block
        line, col = torch.div(body, width + 1, rounding_mode="floor"), body % (width + 1)
        bidx = line * width + col            # base index inside the contig
        L = torch.from_numpy(lens[c0:c1]).to(dev)[cid]
        is_nl = (body >= 0) & ((col == width) | (bidx >= L))
        src = (torch.from_numpy(qoff[c0:c1]).to(dev)[cid] + 1 + bidx).clamp_(0, codes.numel() - 1)
        ch = ascii_lut[codes[src].long()]
        ch = torch.where(is_nl, torch.full_like(ch, 10), ch)
        ch = torch.where(body < 0, hdr_d[c0:c1][cid, off.clamp(max=H - 1)], ch)
        host[int(toff[c0]):int(toff[c1])].copy_(ch)
        del cid, off, body, line, col, bidx, L, is_nl, src, ch
        c0 = c1
    return host


def pack_codes(codes: torch.Tensor):
    """uint8 codes on the device -> (seq int64[W], inv int32[W], n) with the library's device packer."""
    n = codes.numel()
    W = hs.packed_words(n)
    seq = torch.empty(W, dtype=torch.int64, device=codes.device)
    inv = torch.empty(W, dtype=torch.int32, device=codes.device)
    hs.pack_codes_device(codes.data_ptr(), n, seq.data_ptr(), inv.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().synchronize()
    return seq, inv, n


@dataclass
class SketchSet:
    """A synthetic sketch database + the genomes its "real" sketches came from (device codes)."""
    k: int
    s: int
    offsets: np.ndarray
    hashes: np.ndarray
    lengths: np.ndarray
    n_real: int
    genome_len: int
    genomes: torch.Tensor        # uint8 codes [n_real * genome_len] on the device
    seed: int


def make_db(device: int, n_sketches: int = 50_000, n_real: int = 500, genome_len: int = 2_000_000, k: int = 21,
            s: int = 1000, seed: int = 2, cluster_copies: int = 0, tiny: int = 0) -> SketchSet:
    """`n_real` genomes sketched from their sequence with the GPU sketcher + decoy sketches (bottom-s of
    uniform hashes: statistically what sketches of unrelated genomes look like)."""
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    hs._abi.init(device)
    n_real = min(n_real, n_sketches)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)                      # genomes + db identical on every rank
    genomes = torch.randint(0, 4, (n_real * genome_len,), dtype=torch.uint8, device=dev, generator=gen)
    if cluster_copies:
        # config 4: the last `cluster_copies` genomes become near-identical relatives (1-5 % substitutions,
        # mutationGCF.py model) of the first ones -> clusters that winner-take-all has to separate
        cluster_copies = min(cluster_copies, n_real // 2)
        for c in range(cluster_copies):
            src = genomes[(c % (n_real - cluster_copies)) * genome_len:][:genome_len]
            m = 0.01 + 0.04 * ((c * 7919) % 1000) / 1000.0
            hit = torch.rand(genome_len, device=dev, generator=gen) < m
            shift = torch.randint(1, 4, (genome_len,), device=dev, generator=gen, dtype=torch.uint8)
            genomes[(n_real - cluster_copies + c) * genome_len:][:genome_len] = torch.where(hit, (src + shift) % 4, src)
    real = np.zeros((n_real, s), np.uint64)
    for i in range(n_real):
        gseq, ginv, gn = pack_codes(genomes[i * genome_len:(i + 1) * genome_len])
        h = hs.sketch_packed_device(gseq.data_ptr(), ginv.data_ptr(), gn, k, s)
        assert len(h) == s
        real[i] = h
    rng = np.random.default_rng(seed)
    decoy, dlen = synth.decoy_sketches(rng, n_sketches - n_real, s)
    if tiny:   # sketches of tiny genomes: bottom-s of only 1.5-20 thousand k-mers, i.e. spread over the hash range
        tiny = min(tiny, len(decoy))
        decoy[:tiny], dlen[:tiny] = synth.decoy_sketches(rng, tiny, s, g_lo=1.5 * s, g_hi=20.0 * s)
    hashes = np.concatenate([real.reshape(-1), decoy.reshape(-1)])
    real_len = np.full(n_real, genome_len, np.uint64)
    if cluster_copies:  # distinct lengths inside a cluster so (score, length) ties are rare but present
        real_len[n_real - cluster_copies:] -= (np.arange(cluster_copies, dtype=np.uint64) % np.uint64(7))
    lengths = np.concatenate([real_len, dlen])
    offsets = np.arange(n_sketches + 1, dtype=np.uint64) * np.uint64(s)
    return SketchSet(k=k, s=s, offsets=offsets, hashes=hashes, lengths=lengths, n_real=n_real, genome_len=genome_len,
                     genomes=genomes, seed=seed)


def make_query(sk: SketchSet, mbp: int, rate: float = 0.01, shard: int = 0, with_fasta: bool = False,
               with_host_packed: bool = False, fasta_width: int = 80, fasta_sample_mbp: int = 0, name: str = "c2") -> Workload:
    """`mbp` Mbp of contigs cut from the real genomes (Zymo-fitted lengths, `rate` substitutions, half
    reverse-complemented), packed on the device.  Shard `shard` is an independent draw: N ranks with
    shards 0..N-1 hold N disjoint parts of one contig set.  fasta_sample_mbp > 0: also the FASTA text
    of the first contigs up to that many Mbp (host bytes; the CPU oracle's bounded sample)."""
    dev = sk.genomes.device
    qrng = np.random.default_rng(sk.seed * 1000 + shard)      # each rank cuts its own shard of contigs
    gen = torch.Generator(device=dev)
    gen.manual_seed(sk.seed * 1000 + shard)
    lens, g, st, rc = _contig_table(qrng, sk.n_real, sk.genome_len, mbp * 1_000_000)
    codes = _fill_query_codes(sk.genomes, sk.genome_len, lens, g, st, rc, rate, gen)
    d_seq, d_inv, n_pos = pack_codes(codes)
    wl = Workload(name=name, k=sk.k, s=sk.s, offsets=sk.offsets, hashes=sk.hashes, lengths=sk.lengths, n_real=sk.n_real,
                  d_seq=d_seq, d_inv=d_inv, n_positions=n_pos, n_bases=int(lens.sum()), n_contigs=len(lens))
    if with_fasta:
        wl.fasta = _fasta_from_codes(codes, lens, fasta_width)
    if fasta_sample_mbp:
        c = int(np.searchsorted(np.cumsum(lens), fasta_sample_mbp * 1_000_000, side="left")) + 1
        c = min(c, len(lens))
        n_codes = int((lens[:c] + 1).sum())
        wl.fasta_sample = _fasta_from_codes(codes[:n_codes], lens[:c], fasta_width).numpy().tobytes()
    if with_host_packed:
        words = (n_pos + 31) // 32
        wl.h_seq = torch.empty(words, dtype=torch.int64, pin_memory=True).copy_(d_seq[:words])
        wl.h_inv = torch.empty(words, dtype=torch.int32, pin_memory=True).copy_(d_inv[:words])
    del codes
    torch.cuda.empty_cache()
    return wl


def make_c2(device: int, mbp: int = 1000, n_sketches: int = 50_000, n_real: int = 500, genome_len: int = 2_000_000,
            k: int = 21, s: int = 1000, rate: float = 0.01, seed: int = 2, shard: int = 0, with_fasta: bool = True,
            with_host_packed: bool = True, fasta_width: int = 80, cluster_copies: int = 0, tiny: int = 0) -> Workload:
    sk = make_db(device, n_sketches, n_real, genome_len, k, s, seed, cluster_copies, tiny)
    wl = make_query(sk, mbp, rate, shard, with_fasta, with_host_packed, fasta_width)
    del sk
    torch.cuda.empty_cache()
    return wl
