// Internal declarations shared by layout.cpp (pure C++) and idash_b200.cu (CUDA).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "idash_b200.h"
#include "idash_b200_layout.h"

namespace idash_b200 {

// thread-local last-error text behind idash_b200_last_error()
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();

}  // namespace idash_b200

struct idash_b200_layout {
    uint32_t S = 0, NR = 0, RS = 0;
    uint64_t n_rows = 0, nnz = 0;
    std::vector<idash_b200_group> groups;
    std::vector<idash_b200_entry> entries;
    std::vector<uint32_t> out_bidx;   // per caller row
    std::vector<uint64_t> var_ptr;    // per caller row (+1)
    std::vector<uint32_t> var_ct;
    std::vector<double> var_w;
    // band tiles (tensor-core kernel); empty when the model is not eligible
    std::vector<idash_b200_tile> tiles;
    std::vector<uint32_t> tile_rows;
    std::vector<int32_t> tile_bias;
    std::vector<uint8_t> tile_coef;
    std::vector<uint32_t> tile_used;
    uint32_t tile_kmax = 0;           // widest band over all tiles
    bool ring_ok = false;             // eligible for the persistent ring kernel
    std::vector<uint32_t> feat_used;  // ring_ok only: bit f = some row uses input feature f
    uint32_t ct_min = 1, ct_max = 0;
    uint32_t max_entries_per_group = 0;
    bool shifts_aligned = true;
};
