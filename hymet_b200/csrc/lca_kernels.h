// lca_kernels.h -- launch interface of the weighted-LCA kernel (internal to libhymet_screen.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hs {

constexpr uint32_t kLcaRanks = 8;   // superkingdom .. strain (classification_cami.py:16)

struct LcaArgs {
    uint64_t n_q;
    const uint64_t *q_off;   // n_q + 1: alignments of query q are [q_off[q], q_off[q+1])
    const int32_t *tax;      // per alignment: row of `names` for the target's taxid, -1 = no taxid known
    const double *w;         // per alignment: coverage x reference abundance
    const uint32_t *names;   // n_tax x kLcaRanks name ids, 0 = no name at that rank
    int32_t *s_tax;          // scratch, one slot per alignment each
    double *s_w;
    uint32_t *s_name;
    double *s_nw;
    uint32_t *out_names;     // n_q x kLcaRanks: the chosen name per rank, out_depth of them
    uint32_t *out_depth;     // ranks resolved; 0 = "Unknown"
    double *out_conf;
    uint8_t *out_any;        // some alignment had a taxid
};

cudaError_t launch_weighted_lca(const LcaArgs &a, int sm_count, cudaStream_t st);

}  // namespace hs
