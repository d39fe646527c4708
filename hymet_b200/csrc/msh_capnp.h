// msh_capnp.h -- Mash sketch file (.msh) loader: Cap'n Proto stream framing and the
// MinHash schema decoded by hand (no libcapnp in this image).
// Stands for Sketch::initFromFiles of `mash screen` (row a4), whose input is the
// data/sketch1-3.msh files of /root/reference/run_hymet_cami.sh:52,85-96.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace hs {

struct MshData {
    uint32_t k = 0, s = 0, seed = 42;
    bool use64 = true;
    std::vector<std::string> names, comments;
    std::vector<uint64_t> lengths;  // S18: length64 if non-zero else length
    std::vector<uint64_t> offsets;  // n_refs + 1
    std::vector<uint64_t> hashes;   // ascending per reference; 32-bit sketches widened
    double t_parse_s = 0;
};

// Returns 0 or a negative HS_E* code with a message in `err`.
int msh_read(const std::string &path, MshData &out, std::string &err);

}  // namespace hs
