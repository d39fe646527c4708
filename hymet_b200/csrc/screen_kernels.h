// screen_kernels.h -- launch interface of the hand-written sm_100a kernels.
// Internal to libhymet_screen.so (the public surface is include/hymet_screen.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_core.cuh"

namespace hs {

constexpr int kTileWords = 256;    // one CTA pass = 256 words = 8192 bases
constexpr int kCtaThreads = 256;
constexpr uint32_t kNoEntry = 0xFFFFFFFFu;

// counters accumulated by the streaming kernel (device, 64-bit each)
enum StatSlot { ST_VALID = 0, ST_PROBES, ST_BUCKETS, ST_HITS, ST_MIXINS, ST_COUNT };

struct TableView {
    const uint64_t *keys;   // n_buckets * 4, kEmptyKey = free
    const uint32_t *vals;   // n_buckets * 4, canonical entry id of the key
    uint32_t n_buckets;
    uint64_t max_key;       // largest stored key (range pre-filter)
    uint32_t special;       // canonical entry id of key == kEmptyKey, or kNoEntry
    // Second-level pre-filter for databases whose keys are spread over the whole hash range
    // (sketches of tiny genomes keep ALL their k-mer hashes, so max_key ~ 2^64 and the range
    // test stops helping): a blocked Bloom filter, one 64-bit word and 3 bits per key, sized to
    // stay L2 resident.  No false negatives => results unchanged.  NULL when not built.
    const unsigned long long *bloom;
    uint32_t bloom_mask;    // words - 1 (power of two)
};

// Mixture bottom-s state lives on the device so that streaming never waits for the host:
// the threshold is lowered by a maintenance kernel after each chunk, not by a host check.
struct MixState {
    unsigned long long tau;  // accept h <= tau (min over launches of their caps, /4 per shrink)
    uint32_t count;          // distinct values in sets[cur]
    uint32_t new_count;      // scratch while rebuilding into sets[cur ^ 1]
    uint32_t overflow;       // an insert was dropped: the set is incomplete below tau
    uint32_t has_max;        // hash == kEmptyKey seen (only possible while tau == 2^64-1)
    uint32_t cur;            // live set
    uint32_t n_out, n_unique;  // finalisation scratch
    uint32_t rebuilds;
};
struct MixView {
    uint64_t *sets[2];      // open addressing, capacity mask+1, kEmptyKey = free
    uint32_t mask;
    uint32_t limit;         // max distinct before `overflow` is raised
    uint64_t tau_cap;       // this launch's cap: expected offers stay <= capacity/8
    MixState *st;
};

struct StreamArgs {
    const uint64_t *seq;    // packed words (16-byte aligned), n_tiles * kTileWords allocated
    const uint32_t *inv;
    uint64_t n_bases;       // positions >= n_bases are invalid whatever inv says
    const unsigned long long *n_bases_dev;  // if set, the length is read from HBM (device-parsed chunks)
    uint32_t n_tiles;       // END tile (exclusive) of this launch
    uint32_t tile_begin;    // first tile of this launch (> 0: a later piece of the same chunk, halo is real data)
    int k;
    uint32_t seed;
    int use64;
    int do_count;           // probe + count
    int do_filter;          // skip probes for h > max_key
    int do_mix;             // offer hashes <= tau to the mixture set
    TableView tab;
    uint32_t *counts;       // per canonical entry id
    MixView mix;
    unsigned long long *stats;  // ST_COUNT slots
    uint64_t *emit_hash;    // optional: per position hash (K1 parity), n_bases entries
    uint8_t *emit_valid;
};

// ---- streaming (K1+K2+K3 insert) -------------------------------------------
cudaError_t launch_stream(const StreamArgs &a, int sm_count, cudaStream_t st);

// ---- database build (row a5) -------------------------------------------------
cudaError_t launch_table_insert(uint64_t *keys, uint32_t *vals, uint32_t n_buckets, const uint64_t *hashes,
                                uint64_t n_entries, uint32_t *special, unsigned long long *max_key,
                                uint32_t *fail, cudaStream_t st);
cudaError_t launch_table_canon(const TableView &t, const uint64_t *hashes, uint64_t n_entries, uint32_t *canon,
                               unsigned long long *n_distinct, cudaStream_t st);
cudaError_t launch_bloom_build(unsigned long long *bloom, uint32_t bloom_mask, const uint64_t *hashes, uint64_t n,
                               cudaStream_t st);
// standalone probe (K2): out_entry may be NULL; stats[0]=hits, stats[1]=bucket reads
cudaError_t launch_probe(const TableView &t, const uint64_t *hashes, uint64_t n, uint32_t *out_entry,
                         unsigned long long *stats, int sm_count, cudaStream_t st);

// random 32-byte sector reads over `bytes` of device memory (measurement reference for K2)
cudaError_t launch_gather_bench(const void *buf, uint64_t bytes, uint64_t total_reads, int sm_count, cudaStream_t st);

// ---- mixture bottom-s (K3) ---------------------------------------------------
// copy keys <= thr from the set into out (append with atomic cursor); counts all <= thr
cudaError_t launch_mix_collect(const uint64_t *set, uint32_t cap, uint64_t thr, uint64_t *out, uint32_t out_cap,
                               uint32_t *n_out, cudaStream_t st);
// after a streaming launch: fold its cap into tau and, if the live set is more than a
// quarter full, rebuild it at tau/4 into the other buffer (exact: >= s smaller values stay)
cudaError_t launch_mix_maintain(const MixView &v, cudaStream_t st);
// sort ascending + unique in place (n <= cap_pow2 handled by padding); *n_unique out
cudaError_t launch_sort_unique(uint64_t *data, uint32_t n, uint64_t *scratch, uint32_t *n_unique, cudaStream_t st);

// ---- per-sketch reduction (K4), winner-take-all (K5), statistics (K6) ---------
cudaError_t launch_sketch_reduce(const uint64_t *offsets, uint64_t n_refs, const uint32_t *canon,
                                 const uint32_t *counts, const uint32_t *winner /*nullable*/, uint32_t *shared,
                                 uint32_t *median, int sm_count, cudaStream_t st);
cudaError_t launch_winner(const uint64_t *offsets, uint64_t n_refs, const uint32_t *canon, const uint32_t *counts,
                          const uint32_t *shared, const uint64_t *lengths, unsigned long long *best_score,
                          unsigned long long *best_len, uint32_t *winner, uint64_t n_entries, int sm_count,
                          cudaStream_t st);
cudaError_t launch_stats(uint32_t k, uint64_t set_size, uint64_t n, const uint32_t *shared32,
                         const uint64_t *shared64, const uint64_t *offsets, const uint64_t *sizes,
                         double *identity, double *pvalue, cudaStream_t st);

// ---- sparse count exchange (multi-GPU) -------------------------------------------
cudaError_t launch_counts_compact(const uint32_t *counts, uint64_t n, unsigned long long *pairs, uint32_t cap,
                                  uint32_t *n_out, cudaStream_t st);
cudaError_t launch_counts_scatter_add(uint32_t *counts, uint64_t n_counts, const unsigned long long *pairs,
                                      uint64_t n_pairs, cudaStream_t st);

// ---- device-side packer (codes -> 2-bit words + invalid mask) -------------------
cudaError_t launch_pack_codes(const uint8_t *codes, uint64_t n, uint64_t *seq, uint32_t *inv, uint64_t n_words_alloc,
                              cudaStream_t st);

}  // namespace hs

// ---- device-side FASTA ingest (row a6 on the GPU) --------------------------------
// Raw FASTA bytes in HBM -> base codes (0..3, 4 = invalid / record separator), compacted:
// headers and line ends removed, one separator per record, exactly like fasta_pack.cpp.
namespace hs {
constexpr int kFaTileBytes = 8192;   // 256 threads x 32 bytes
struct FaScratch {
    uint32_t *tile_event;   // per tile: its last newline/header-start event code (0 = none)
    uint32_t *tile_carry;   // exclusive prefix max of tile_event
    uint32_t *tile_event_b; // per tile: its last header-start/'+'-line event code
    uint32_t *tile_carry_b; // exclusive prefix max of tile_event_b, seeded "skipping"
    uint4 *masks;           // per 32 input bytes: newline / header-start / '+'-start / dropped-CR masks
    uint2 *emit;            // per 32 input bytes: emitted-byte mask, separator mask
    uint32_t *tile_count;   // per tile: positions emitted
    uint64_t *tile_offset;  // exclusive prefix sum of tile_count
    unsigned long long *totals;  // [0] positions, [1] sequence bases, [2] records  (accumulated)
    unsigned long long *chunk_positions;  // positions of THIS chunk (read by the pack/stream launches)
};
// text: n bytes (n < 2^30) starting at a record header; codes: capacity >= n
cudaError_t launch_fasta_to_codes(const uint8_t *text, uint32_t n, uint8_t *codes, const FaScratch &sc, cudaStream_t st);
// k_pack_codes / k_stream variants whose length comes from device memory (no host sync)
cudaError_t launch_pack_codes_dyn(const uint8_t *codes, const unsigned long long *n_dev, uint64_t *seq, uint32_t *inv,
                                  uint64_t n_words_alloc, cudaStream_t st);
}  // namespace hs
