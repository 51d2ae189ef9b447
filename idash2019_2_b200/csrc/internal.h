// Internal declarations shared by layout.cpp (pure C++) and idash_b200.cu (CUDA).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <memory>
#include <utility>
#include <vector>

#include "idash_b200.h"
#include "idash_b200_layout.h"

namespace idash_b200 {

// thread-local last-error text behind idash_b200_last_error()
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();

// worker threads of the host-side passes (IDASH_B200_THREADS, default: hardware concurrency, at most 64)
unsigned host_threads();

// std::vector whose resize(n) leaves new elements uninitialised: the big arrays of the layout are written exactly once by the
// parallel passes, and a serial zero fill of ~100 MB would cost more than the passes themselves. Large arrays come from
// huge-page mappings: first-touch faults of 4 KB pages are serialised by the kernel's address-space lock and were the whole
// cost of the passes (no speed-up from threads at all).
void *big_alloc(size_t bytes);           // >= 4 MB: anonymous mapping with transparent huge pages; else malloc
void big_free(void *p, size_t bytes);
}  // namespace idash_b200

template <class T>
struct default_init_allocator {
    using value_type = T;
    default_init_allocator() noexcept = default;
    template <class U> default_init_allocator(const default_init_allocator<U> &) noexcept {}
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    T *allocate(size_t n) { return static_cast<T *>(idash_b200::big_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t n) noexcept { idash_b200::big_free(p, n * sizeof(T)); }
    template <class U, class... Args>
    void construct(U *p, Args &&...args) {
        if constexpr (sizeof...(args) == 0) ::new (static_cast<void *>(p)) U;
        else ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
    template <class U> bool operator==(const default_init_allocator<U> &) const noexcept { return true; }
    template <class U> bool operator!=(const default_init_allocator<U> &) const noexcept { return false; }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

struct idash_b200_layout {
    std::atomic<int> refs{1};         // models on several devices share one layout (idash_b200_model_clone)
    uint32_t S = 0, NR = 0, RS = 0;
    uint64_t n_rows = 0, nnz = 0;
    std::vector<uint32_t> out_bidx;   // per caller row
    // the model itself, normalised: per caller row the non-constant entries sorted by input bigIndex + the Constant
    std::vector<uint64_t> feat_ptr;   // per caller row (+1)
    raw_vector<uint32_t> feat_bidx;
    raw_vector<int32_t> feat_coef;
    std::vector<int32_t> bias;        // per caller row
    std::vector<uint32_t> order;      // caller rows sorted by output bigIndex
    // variance CSR per caller row, entries in the CALLER's order (region 0 only)
    std::vector<uint64_t> var_ptr;
    raw_vector<uint32_t> var_ct;
    raw_vector<double> var_w;
    raw_vector<double> var_wsum;     // per caller row: sum of its var_w
    // band tiles (tensor-core kernels); empty when no row is eligible
    std::vector<idash_b200_tile> tiles;
    raw_vector<uint32_t> tile_rows;
    raw_vector<int32_t> tile_bias;
    raw_vector<uint8_t> tile_coef;
    raw_vector<uint32_t> tile_used;
    uint32_t tile_kmax = 0;           // widest band over all tiles
    bool ring_ok = false;             // eligible for the persistent ring kernel
    std::vector<uint32_t> feat_used;  // ring_ok only: bit f = some tile row uses input feature f
    // IMAD groups of the OVERFLOW rows (rows no tile holds) ...
    std::vector<idash_b200_group> groups;
    std::vector<idash_b200_entry> entries;
    // ... and, once groups_all is set (COMPILE_GROUPS_ALL / layout_ensure_groups_all), of every row
    std::vector<idash_b200_group> groups_full;
    std::vector<idash_b200_entry> entries_full;
    bool groups_all = false;
    uint64_t n_overflow_rows = 0;     // rows that no tile holds (a coefficient outside the limb range, or a band wider than a tile)
    uint32_t ct_min = 1, ct_max = 0;
    uint32_t max_entries_per_group = 0;
    bool shifts_aligned = true;
};
