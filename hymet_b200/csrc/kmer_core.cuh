// kmer_core.cuh -- arithmetic shared by every kernel on the screen hot path.
//
// Restates, for 2-bit packed sequence, what `mash screen` does per k-mer
// (the binary HYMET runs at /root/reference/scripts/mash.sh:14; rules S1-S5 of
// SURVEY.md Appendix A): canonical k-mer = min(fwd, revcomp) byte-wise,
// hash = MurmurHash3_x64_128(ASCII k-mer, seed)[h1], low 32 bits when 4^k <= 2^32.
//
// Everything here is `HS_HD` (host+device) on purpose: tests/ compiles this
// header with g++ and checks each function against the CPU oracle before any
// GPU time is spent.  The product only ever calls it from CUDA kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

namespace hs {

// ---- packed query layout --------------------------------------------------
// seq word w holds bases 32w..32w+31, base j at bits [62-2j, 63-2j] (first base
// most significant), codes A=0 C=1 G=2 T=3.  inv word w: bit (31-j) set when
// base j is not A/C/G/T (N, IUPAC, record separator, padding).
constexpr int kBasesPerWord = 32;

HS_HD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// Reverse the order of the 32 two-bit groups of x.
HS_HD uint64_t pair_reverse64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
#endif
    // full bit reversal also swapped the two bits inside every pair: swap back
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// For the 64 bases (prev word, cur word): bit (31-j) of the result is set when
// the k-mer ENDING at base j of the current word covers an invalid base.
HS_HD uint32_t invalid_kmer_ends(uint32_t inv_prev, uint32_t inv_cur, int k)
{
    uint64_t acc = ((uint64_t)inv_prev << 32) | inv_cur;
    int w = 1;
    while (2 * w <= k) { acc |= acc >> w; w *= 2; }  // bit p = OR of V[p .. p+w-1]
    if (w < k) acc |= acc >> (k - w);                // ... V[p .. p+k-1]
    return (uint32_t)acc;
}

// Rolling state of one thread: the k-mer ending at the most recent base.
//   fm  : forward k-mer, first base most significant, right aligned, masked
//   flp : forward k-mer, first base LEAST significant, LEFT aligned in 64 bits
//         (bits below 64-2k hold older bases and are ignored)
struct Roll { uint64_t fm, flp; };

HS_HD Roll roll_init(uint64_t prev_word, int k)
{
    Roll r;
    r.fm = prev_word & kmer_mask(k);     // k-mer ending at base 31 of the previous word
    r.flp = pair_reverse64(r.fm);
    return r;
}

HS_HD void roll_push(Roll &r, uint32_t code, int k)
{
    r.fm = ((r.fm << 2) | code) & kmer_mask(k);
    r.flp = (r.flp >> 2) | ((uint64_t)code << 62);
}

// S5: canonical k-mer, returned first-base-least-significant (byte order of the
// ASCII string murmur will read).  fwd <= rc compared first-base-most-significant
// == memcmp of the ASCII strings because 'A'<'C'<'G'<'T' and codes are 0..3.
HS_HD uint64_t canonical_lsb(const Roll &r, int k)
{
    const uint64_t mask = kmer_mask(k);
    const uint64_t fl = r.flp >> (64 - 2 * k);  // forward, LSB-first
    const uint64_t rm = (~fl) & mask;           // reverse complement, MSB-first
    return (r.fm <= rm) ? fl : ((~r.fm) & mask);
}

// ---- 2-bit -> ASCII expansion ---------------------------------------------
// ascii4(b): the four ASCII letters of the 4 bases in byte b (base i at bits
// [2i,2i+1]), first base in the low byte: what a little-endian load of the
// k-mer string would see.
HS_HD uint32_t ascii4(uint32_t b)
{
    const uint32_t lut = 0x54474341u;  // 'A','C','G','T' for codes 0..3
    return ((lut >> (8 * (b & 3))) & 0xFFu) | (((lut >> (8 * ((b >> 2) & 3))) & 0xFFu) << 8) |
           (((lut >> (8 * ((b >> 4) & 3))) & 0xFFu) << 16) | (((lut >> (8 * ((b >> 6) & 3))) & 0xFFu) << 24);
}

// ---- MurmurHash3_x64_128 ----------------------------------------------------
HS_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
HS_HD uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// h1 of MurmurHash3_x64_128 over the k (<= 32) ASCII bytes given as eight
// little-endian 32-bit words w[0..7]; bytes at index >= k MUST be zero.
HS_HD uint64_t murmur3_h1_words(const uint32_t w[8], int k, uint32_t seed)
{
    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k >> 4;
    int base = 0;
    for (int b = 0; b < 2; b++) {
        if (b < nblocks) {
            uint64_t k1 = (uint64_t)w[base] | ((uint64_t)w[base + 1] << 32);
            uint64_t k2 = (uint64_t)w[base + 2] | ((uint64_t)w[base + 3] << 32);
            k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
            h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
            h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
            base += 4;
        }
    }
    if (k & 15) {
        // tail: zero padding makes the unconditional k2 step a no-op when <= 8 bytes remain
        uint64_t k1 = (uint64_t)w[base & 7] | ((uint64_t)w[(base + 1) & 7] << 32);
        uint64_t k2 = ((k & 15) > 8) ? ((uint64_t)w[(base + 2) & 7] | ((uint64_t)w[(base + 3) & 7] << 32)) : 0ull;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// Expand a canonical LSB-first k-mer and hash it.  `Lut(cl, i)` returns the 4 ASCII
// letters of byte i of cl: a shared-memory table on the device, plain arithmetic on the host.
template <class Lut>
HS_HD uint64_t hash_canonical(uint64_t cl, int k, uint32_t seed, bool use64, const Lut &lut)
{
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int nb = k - 4 * i;  // bases available for this word
        uint32_t v = 0;
        if (nb > 0) {
            v = lut(cl, i);  // letters of bases 4i..4i+3
            if (nb < 4) v &= (1u << (8 * nb)) - 1u;
        }
        w[i] = v;
    }
    const uint64_t h = murmur3_h1_words(w, k, seed);
    return use64 ? h : (h & 0xFFFFFFFFull);
}

struct AsciiArith {
    HS_HD uint32_t operator()(uint64_t cl, int i) const { return ascii4((uint32_t)(cl >> (8 * i)) & 0xFFu); }
};

// One thread's unit of work: the 32 k-mers that END inside word `cur` (their
// first bases may lie in `prev`; k <= 32 so never further back).  sink(j, hash)
// is called for every valid k-mer ending at base j of `cur`, in order.
template <class Lut, class Sink>
HS_HD void for_each_kmer_in_word(uint64_t prev, uint64_t cur, uint32_t inv_prev, uint32_t inv_cur, int k,
                                 uint32_t seed, bool use64, const Lut &lut, Sink &&sink)
{
    const uint32_t bad = invalid_kmer_ends(inv_prev, inv_cur, k);
    Roll r = roll_init(prev, k);
#pragma unroll 4
    for (int j = 0; j < kBasesPerWord; j++) {
        roll_push(r, (uint32_t)(cur >> (62 - 2 * j)) & 3u, k);
        if (!((bad >> (31 - j)) & 1u)) sink(j, hash_canonical(canonical_lsb(r, k), k, seed, use64, lut));
    }
}

// ---- sketch hash table ------------------------------------------------------
// Buckets of four 8-byte keys (one 32-byte DRAM sector); kEmpty marks a free
// slot.  Reference hashes are bottom-s values (numerically small) so the bucket
// index re-mixes both halves before the multiply-high range reduction.
constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kBucketSlots = 4;

HS_HD uint32_t bucket_of(uint64_t h, uint32_t n_buckets)
{
    uint32_t m = (uint32_t)h ^ ((uint32_t)(h >> 32) * 0x9E3779B1u);
    m *= 0x85EBCA77u;
    m ^= m >> 15;
    m *= 0xC2B2AE3Du;
#if defined(__CUDA_ARCH__)
    return __umulhi(m, n_buckets);
#else
    return (uint32_t)(((uint64_t)m * n_buckets) >> 32);
#endif
}

// Blocked Bloom filter: which 64-bit word, and which 3 bits inside it, belong to hash h.
HS_HD void bloom_slot(uint64_t h, uint32_t word_mask, uint32_t &word, unsigned long long &bits)
{
    const uint64_t g = h * 0x9E3779B97F4A7C15ull;
    word = (uint32_t)(g >> 34) & word_mask;
    bits = (1ull << (g & 63)) | (1ull << ((g >> 6) & 63)) | (1ull << ((g >> 12) & 63));
}

HS_HD uint32_t mixset_slot(uint64_t h, uint32_t mask)
{
    uint64_t x = h * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(x >> 40) & mask;
}

}  // namespace hs
