// tests/host_emul/host_emul.cpp -- TEST INFRASTRUCTURE.
// Compiles the product's host/device-shared arithmetic (csrc/kmer_core.cuh) and the
// product's FASTA packer with g++ and walks the packed words exactly the way one CUDA
// thread per word does, so the layout / rolling / canonical / murmur arithmetic is
// checked against the oracle on the CPU box before GPU time is spent.  Nothing in
// hymet_b200/ links this file; it is not a CPU fallback.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../hymet_b200/csrc/fasta_pack.h"
#include "../../hymet_b200/csrc/kmer_core.cuh"

extern "C" {

// returns number of hashes written (valid k-mers, stream order); stats[0..2] = records, bases, positions
int64_t emul_pack_and_hash(const char *text, uint64_t n, int k, uint32_t seed, uint64_t *out, uint64_t cap,
                           uint64_t *stats)
{
    std::vector<uint64_t> seq(hs::pack_words_bound(n));
    std::vector<uint32_t> inv(hs::pack_words_bound(n));
    hs::PackStats st;
    uint64_t nw = hs::pack_text_span(text, n, seq.data(), inv.data(), &st);
    stats[0] = st.n_records; stats[1] = st.n_seq_bases; stats[2] = st.n_positions;
    const bool use64 = k > 16;
    uint64_t m = 0;
    bool overflow = false;
    for (uint64_t w = 0; w < nw; w++) {
        const uint64_t prev = w ? seq[w - 1] : 0;
        const uint32_t iprev = w ? inv[w - 1] : ~0u;
        hs::for_each_kmer_in_word(prev, seq[w], iprev, inv[w], k, seed, use64, hs::AsciiArith(),
                                  [&](int, uint64_t h) { if (m < cap) out[m++] = h; else overflow = true; });
    }
    return overflow ? -1 : (int64_t)m;
}

uint64_t emul_pack(const char *text, uint64_t n, uint64_t *seq, uint32_t *inv, uint64_t *stats)
{
    hs::PackStats st;
    uint64_t nw = hs::pack_text_span(text, n, seq, inv, &st);
    stats[0] = st.n_records; stats[1] = st.n_seq_bases; stats[2] = st.n_positions;
    return nw;
}

// both hash formulations on a raw LSB-first k-mer value (bits above 2k are masked off)
void emul_hash_both(uint64_t cl, int k, uint32_t seed, uint64_t *out)
{
    cl &= hs::kmer_mask(k);
    out[0] = hs::hash_canonical(cl, k, seed, k > 16, hs::AsciiArith());
    out[1] = hs::hash_canonical_premul(cl, k, seed, k > 16, hs::PremulArith());
}

// the 32 k-mers ending in `cur`: rolling formulation vs windowed top-aligned k-mers + pre-multiplied tables.
// out_a/out_b[j] = hash (validity is a separate mask and not involved); returns mismatches.
int emul_window_vs_roll(uint64_t prev, uint64_t cur, int k, uint32_t seed, uint64_t *out_a, uint64_t *out_b)
{
    const bool use64 = k > 16;
    hs::Roll r = hs::roll_init(prev, k);
    const hs::Win wt = hs::win_init_top(prev, cur, k);
    int bad = 0;
    for (int j = 0; j < 32; j++) {
        hs::roll_push(r, (uint32_t)(cur >> (62 - 2 * j)) & 3u, k);
        out_a[j] = hs::hash_canonical(hs::canonical_lsb(r, k), k, seed, use64, hs::AsciiArith());
        out_b[j] = hs::hash_canonical_premul_top(hs::canonical_top(wt, j), k, seed, use64, hs::PremulArithMsb());
        bad += out_a[j] != out_b[j];
    }
    return bad;
}

uint32_t emul_bucket_of(uint64_t h, uint32_t nb) { return hs::bucket_of(h, nb); }
uint64_t emul_pair_reverse(uint64_t x) { return hs::pair_reverse64(x); }

int emul_split(const char *text, uint64_t n, int parts, uint64_t min_span, uint64_t *bounds, int cap)
{
    auto v = hs::split_records(text, n, parts, min_span);
    int m = 0;
    for (auto &p : v) { if (m < cap) { bounds[2 * m] = p.first; bounds[2 * m + 1] = p.second; } m++; }
    return m;
}
}
