// lca_kernels.cu -- SURVEY.md 8f rank 4: the weighted-LCA vote of HYMET's classifier on the GPU.
//
// Path replaced: _process_one + _weighted_lca of /root/reference/scripts/classification_cami.py:251-308.
// For every query (contig) the reference walks its alignments, sums `coverage x reference abundance`
// per taxid, and then, rank by rank from superkingdom down, lets the taxids vote for the name they
// carry at that rank: the heaviest name wins, its share of the voting weight multiplies into the
// confidence, and the walk stops at the first rank where nobody has a name.  Queries are independent
// (the reference spreads them over a process pool): here one thread owns one query.
//
// Bit-exactness.  The reference is CPython: every sum is a chain of IEEE double additions in dict
// insertion order, `max` returns the FIRST maximal item.  The kernel keeps both orders (taxids and
// names in order of first appearance) and uses the _rn intrinsics so that nothing is contracted
// into a fused multiply-add: the confidence comes out bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "lca_kernels.h"

namespace hs {

__global__ void __launch_bounds__(128) k_weighted_lca(const LcaArgs a)
{
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < a.n_q; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t beg = a.q_off[q], end = a.q_off[q + 1];
        int32_t *tid = a.s_tax + beg;      // this query's scratch: distinct taxids in order of first appearance
        double *tw = a.s_w + beg;          // ... and their summed weights
        uint32_t *nid = a.s_name + beg;    // per rank: distinct names in order of first appearance
        double *nw = a.s_nw + beg;
        uint32_t n_t = 0;
        bool any = false;
        // tw[tid] += cov * abundance, in alignment order (classification_cami.py:291-299)
        for (uint64_t j = beg; j < end; j++) {
            const int32_t t = a.tax[j];
            if (t < 0) continue;           // no taxid for this target
            any = true;
            uint32_t p = 0;
            while (p < n_t && tid[p] != t) p++;
            if (p == n_t) { tid[n_t] = t; tw[n_t] = 0.0; n_t++; }
            tw[p] = __dadd_rn(tw[p], a.w[j]);
        }
        double total = 0.0;
        for (uint32_t p = 0; p < n_t; p++) total = __dadd_rn(total, tw[p]);
        uint32_t depth = 0;
        double conf = 1.0;
        if (any && total > 0.0) {
            for (uint32_t r = 0; r < kLcaRanks; r++) {
                uint32_t n_n = 0;
                double denom = 0.0;
                for (uint32_t p = 0; p < n_t; p++) {
                    const uint32_t nm = a.names[(size_t)tid[p] * kLcaRanks + r];
                    if (!nm) continue;     // this taxid has no name at this rank (or no lineage at all)
                    uint32_t x = 0;
                    while (x < n_n && nid[x] != nm) x++;
                    if (x == n_n) { nid[n_n] = nm; nw[n_n] = 0.0; n_n++; }
                    nw[x] = __dadd_rn(nw[x], tw[p]);
                    denom = __dadd_rn(denom, tw[p]);
                }
                if (!(denom > 0.0) || n_n == 0) break;
                uint32_t best = 0;
                for (uint32_t x = 1; x < n_n; x++)
                    if (nw[x] > nw[best]) best = x;          // strict: the first maximal name wins, as Python's max
                a.out_names[q * kLcaRanks + depth] = nid[best];
                conf = __dmul_rn(conf, __ddiv_rn(nw[best], denom));
                depth++;
            }
        }
        a.out_depth[q] = depth;
        a.out_conf[q] = depth ? (conf < 1.0 ? conf : 1.0) : 0.0;
        a.out_any[q] = any ? 1 : 0;
    }
}

cudaError_t launch_weighted_lca(const LcaArgs &a, int sm_count, cudaStream_t st)
{
    if (!a.n_q) return cudaSuccess;
    uint64_t grid = (a.n_q + 127) / 128;
    if (grid > (uint64_t)sm_count * 16) grid = (uint64_t)sm_count * 16;
    k_weighted_lca<<<(uint32_t)grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace hs
