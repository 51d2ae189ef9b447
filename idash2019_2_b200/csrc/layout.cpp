// Model -> device block-banded layout compiler (host, pure C++; no CUDA in this file).
//
// Replaces, for the B200 path, what the reference does on every call of cloud_compute_score:
// deep-copying Model::model into a vector (eval/idash.cpp:772) and walking per-output hash maps of
// (input bigIndex -> coefficient) (eval/idash.cpp:800-819). See include/idash_b200_layout.h for
// the layout and DESIGN.md for why it has this shape.
//
// Rows are independent and tiles are independent once their offsets are known, so every pass below is a parallel loop
// over rows or tiles with a serial prefix sum in between (iDASH scale, 242 646 rows: ~10 ms on 16 threads, was 250 ms).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <chrono>

#include <sys/mman.h>

#include "internal.h"

namespace idash_b200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }

unsigned host_threads() {
    if (const char *e = getenv("IDASH_B200_THREADS")) { const int v = atoi(e); if (v > 0) return (unsigned) std::min(v, 256); }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hc ? hc : 4u, 64u));
}

void *big_alloc(size_t bytes) {
    if (bytes >= ((size_t) 4 << 20)) {
        const size_t huge = (size_t) 2 << 20, len = (bytes + huge - 1) & ~(huge - 1);
        void *p = ::mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) throw std::bad_alloc();
        ::madvise(p, len, MADV_HUGEPAGE);
        return p;
    }
    void *p = malloc(bytes ? bytes : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void big_free(void *p, size_t bytes) {
    if (!p) return;
    if (bytes >= ((size_t) 4 << 20)) {
        const size_t huge = (size_t) 2 << 20;
        ::munmap(p, (bytes + huge - 1) & ~(huge - 1));
        return;
    }
    free(p);
}

namespace {

// fn(begin, end, worker) over [0, n) cut into one contiguous range per worker
template <class F>
void parallel_ranges(uint64_t n, uint64_t grain, F fn) {
    const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>(host_threads(), n / std::max<uint64_t>(grain, 1)));
    if (nt <= 1) { fn((uint64_t) 0, n, 0u); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (uint64_t t = 0; t < nt; ++t) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt, (unsigned) t); });
    for (auto &x : th) x.join();
}

struct Triple {        // the (up to) three variant rows of one target SNP
    uint32_t target;   // out_bidx / 3
    int64_t row[3] = {-1, -1, -1};      // caller row per variant
};

// IMAD groups (two target SNPs x three variants, include/idash_b200_layout.h) of the given caller rows, which must be in
// ascending output-bigIndex order
void build_groups(idash_b200_layout *L, const std::vector<uint32_t> &rows, std::vector<idash_b200_group> &groups,
                  std::vector<idash_b200_entry> &entries) {
    groups.clear();
    entries.clear();
    const uint32_t NR = L->NR, RS = L->RS;
    std::vector<Triple> triples;
    for (uint32_t r : rows) {
        const uint32_t target = L->out_bidx[r] / 3, variant = L->out_bidx[r] % 3;
        if (triples.empty() || triples.back().target != target) {
            Triple t;
            t.target = target;
            triples.push_back(t);
        }
        triples.back().row[variant] = r;
    }
    const uint64_t n_groups = (triples.size() + 1) / 2;
    groups.resize(n_groups);
    struct Acc { uint32_t bidx; int32_t c[6]; uint8_t used; };
    // pass 1 (parallel): the merged entry list of every group; pass 2: offsets; pass 3 (parallel): copy out
    std::vector<std::vector<idash_b200_entry>> per_group(n_groups);
    parallel_ranges(n_groups, 256, [&](uint64_t g0, uint64_t g1, unsigned) {
        std::vector<Acc> uni;
        for (uint64_t g = g0; g < g1; ++g) {
            idash_b200_group &G = groups[g];
            memset(&G, 0, sizeof(G));
            uni.clear();
            for (int half = 0; half < 2; ++half) {
                const uint64_t ti = 2 * g + half;
                for (int v = 0; v < 3; ++v) {
                    const int slot = 3 * half + v;
                    G.row[slot] = IDASH_B200_NO_ROW;
                    if (ti >= triples.size() || triples[ti].row[v] < 0) continue;
                    const uint32_t r = (uint32_t) triples[ti].row[v];
                    G.row[slot] = r;
                    G.bias[slot] = L->bias[r];
                    for (uint64_t e = L->feat_ptr[r]; e < L->feat_ptr[r + 1]; ++e) {
                        if (L->feat_coef[e] == 0) continue;   // contributes nothing to words or variance
                        Acc a;
                        a.bidx = L->feat_bidx[e];
                        memset(a.c, 0, sizeof(a.c));
                        a.c[slot] = L->feat_coef[e];
                        a.used = (uint8_t) (1u << half);
                        uni.push_back(a);
                    }
                }
            }
            std::sort(uni.begin(), uni.end(), [](const Acc &a, const Acc &b) { return a.bidx < b.bidx; });
            size_t m = 0;   // merge equal bidx
            for (size_t i = 0; i < uni.size(); ++i) {
                if (m && uni[m - 1].bidx == uni[i].bidx) {
                    for (int k = 0; k < 6; ++k) uni[m - 1].c[k] += uni[i].c[k];   // disjoint slots: plain merge
                    uni[m - 1].used |= uni[i].used;
                } else {
                    uni[m++] = uni[i];
                }
            }
            uni.resize(m);
            auto &dst = per_group[g];
            for (int cls = 1; cls <= 3; ++cls) {          // used == 1: A only, 3: both, 2: B only
                const uint8_t want = cls == 1 ? 1 : (cls == 2 ? 3 : 2);
                uint32_t cnt = 0;
                for (const Acc &a : uni) {
                    if (a.used != want) continue;
                    idash_b200_entry E;
                    E.ct = a.bidx / NR;
                    E.shift = (a.bidx % NR) * RS;
                    memcpy(E.coef, a.c, sizeof(E.coef));
                    dst.push_back(E);
                    ++cnt;
                }
                if (cls == 1) G.n_a = cnt; else if (cls == 2) G.n_ab = cnt; else G.n_b = cnt;
            }
        }
    });
    uint64_t total = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
        L->max_entries_per_group = std::max<uint32_t>(L->max_entries_per_group, (uint32_t) per_group[g].size());
        groups[g].entry_begin = (uint32_t) total;
        total += per_group[g].size();
    }
    entries.resize(total);
    parallel_ranges(n_groups, 256, [&](uint64_t g0, uint64_t g1, unsigned) {
        for (uint64_t g = g0; g < g1; ++g)
            if (!per_group[g].empty()) memcpy(&entries[groups[g].entry_begin], per_group[g].data(), per_group[g].size() * sizeof(idash_b200_entry));
    });
}

}  // namespace
}  // namespace idash_b200

using namespace idash_b200;

extern "C" const char *idash_b200_last_error(void) { return idash_b200::g_err; }

extern "C" int idash_b200_layout_compile_ex(const idash_b200_model_desc *d, uint32_t flags, idash_b200_layout **out) {
    clear_error();
    if (!d || !out) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: null argument");
    *out = nullptr;
    const uint32_t S = d->num_samples, NR = d->num_regions, RS = d->region_size;
    if (NR == 0 || RS == 0 || (uint64_t) NR * RS > IDASH_B200_N || S > IDASH_B200_N)
        return set_error(IDASH_B200_ERR_INVALID,
                         "layout_compile: bad geometry S=%u NUM_REGIONS=%u REGION_SIZE=%u (need NR*RS <= 1024, S <= 1024)",
                         S, NR, RS);
    const uint64_t n_rows = d->n_rows;
    if (n_rows && (!d->out_bidx || !d->row_ptr)) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: null row arrays");
    if (n_rows >= 0xFFFFFFFFull) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: too many rows");
    const uint64_t nnz = n_rows ? d->row_ptr[n_rows] : 0;
    if (nnz && (!d->col || !d->coef)) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: null entry arrays");
    for (uint64_t r = 0; r < n_rows; ++r)
        if (d->row_ptr[r] > d->row_ptr[r + 1]) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: row_ptr not monotone at row %llu", (unsigned long long) r);
    if (nnz >= (1ull << 40)) return set_error(IDASH_B200_ERR_INVALID, "layout_compile: too many entries");

    idash_b200_layout *L = new (std::nothrow) idash_b200_layout();
    if (!L) return set_error(IDASH_B200_ERR_NOMEM, "layout_compile: out of memory");
    const bool timing = getenv("IDASH_B200_LAYOUT_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[layout_compile] %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    try {
        L->S = S; L->NR = NR; L->RS = RS; L->n_rows = n_rows; L->nnz = nnz;
        L->out_bidx.assign(d->out_bidx, d->out_bidx + n_rows);

        // rows sorted by output bigIndex (genomic order when the targets file is sorted); already sorted is the common case
        L->order.resize(n_rows);
        std::iota(L->order.begin(), L->order.end(), 0u);
        bool ascending = true;
        for (uint64_t i = 1; i < n_rows && ascending; ++i) ascending = d->out_bidx[i - 1] < d->out_bidx[i];
        if (!ascending) {
            std::stable_sort(L->order.begin(), L->order.end(), [&](uint32_t a, uint32_t b) { return d->out_bidx[a] < d->out_bidx[b]; });
            for (uint64_t i = 1; i < n_rows; ++i)
                if (d->out_bidx[L->order[i]] == d->out_bidx[L->order[i - 1]]) {
                    const uint32_t dup = d->out_bidx[L->order[i]];
                    delete L;
                    return set_error(IDASH_B200_ERR_INVALID, "layout_compile: duplicate output bigIndex %u", dup);
                }
        }
        const std::vector<uint32_t> &order = L->order;
        lap("order");

        // ---- per row: Constant, entry counts (features, variance terms)
        L->feat_ptr.assign(n_rows + 1, 0);
        L->var_ptr.assign(n_rows + 1, 0);
        L->bias.assign(n_rows, 0);
        std::atomic<uint64_t> bad_row(UINT64_MAX);
        parallel_ranges(n_rows, 4096, [&](uint64_t r0, uint64_t r1, unsigned) {
            for (uint64_t r = r0; r < r1; ++r) {
                uint64_t nf = 0, nv = 0;
                bool have_const = false;
                for (uint64_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e) {
                    if (d->col[e] == IDASH_B200_CONSTANT_BIDX) {
                        if (have_const) { uint64_t exp = UINT64_MAX; bad_row.compare_exchange_strong(exp, r); }
                        have_const = true;
                        L->bias[r] = d->coef[e];
                    } else {
                        ++nf;
                        if (NR == 1 || d->col[e] % NR == 0) ++nv;
                    }
                }
                L->feat_ptr[r + 1] = nf;
                L->var_ptr[r + 1] = nv;
            }
        });
        if (bad_row != UINT64_MAX) {
            const uint64_t r = bad_row;
            delete L;
            return set_error(IDASH_B200_ERR_INVALID, "layout_compile: row %llu has two Constant entries", (unsigned long long) r);
        }
        lap("count");
        for (uint64_t r = 0; r < n_rows; ++r) { L->feat_ptr[r + 1] += L->feat_ptr[r]; L->var_ptr[r + 1] += L->var_ptr[r]; }
        L->feat_bidx.resize(L->feat_ptr[n_rows]);
        L->feat_coef.resize(L->feat_ptr[n_rows]);
        L->var_ct.resize(L->var_ptr[n_rows]);
        L->var_w.resize(L->var_ptr[n_rows]);
        L->var_wsum.resize(n_rows);

        // ---- per row: entries sorted by input bigIndex; variance terms in the CALLER's entry order (tLweAddMulTo adds them in
        // the order the reference walks the row's map, eval/idash.cpp:800-817); band and coefficient range of the row
        lap("alloc rows");
        struct RowStat { uint32_t fmin, fmax; int32_t cmin, cmax; };
        std::vector<RowStat> stat(n_rows);
        const unsigned n_workers = host_threads();
        struct alignas(64) WorkerStat { uint32_t ct_min = 0xFFFFFFFFu, ct_max = 0; bool has = false; };
        std::vector<WorkerStat> wstat(n_workers + 1);
        std::atomic<uint64_t> dup_row(UINT64_MAX);
        std::atomic<bool> unaligned(false);
        parallel_ranges(n_rows, 2048, [&](uint64_t r0, uint64_t r1, unsigned w) {
            std::vector<std::pair<uint32_t, int32_t>> tmp;
            uint32_t my_ct_min = 0xFFFFFFFFu, my_ct_max = 0;
            bool my_has = false, my_unaligned = false;
            for (uint64_t r = r0; r < r1; ++r) {
                uint64_t k = L->feat_ptr[r], kv = L->var_ptr[r];
                const uint64_t k0 = k;
                bool sorted = true;
                double wsum = 0.;
                for (uint64_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e) {
                    const uint32_t b = d->col[e];
                    if (b == IDASH_B200_CONSTANT_BIDX) continue;
                    if (k > k0 && L->feat_bidx[k - 1] >= b) sorted = false;
                    L->feat_bidx[k] = b;
                    L->feat_coef[k] = d->coef[e];
                    ++k;
                    if (NR == 1 || b % NR == 0) {
                        // tLweAddMulTo: current_variance += (p * p) * sample->current_variance with an int32 p * p
                        const int32_t pp = (int32_t) ((uint32_t) d->coef[e] * (uint32_t) d->coef[e]);
                        L->var_ct[kv] = NR == 1 ? b : b / NR;
                        L->var_w[kv] = (double) pp;
                        wsum += (double) pp;
                        ++kv;
                    }
                }
                L->var_wsum[r] = wsum;
                if (!sorted) {
                    tmp.clear();
                    for (uint64_t i = k0; i < k; ++i) tmp.push_back({L->feat_bidx[i], L->feat_coef[i]});
                    std::sort(tmp.begin(), tmp.end(), [](const std::pair<uint32_t, int32_t> &a, const std::pair<uint32_t, int32_t> &b) { return a.first < b.first; });
                    for (uint64_t i = k0; i < k; ++i) { L->feat_bidx[i] = tmp[i - k0].first; L->feat_coef[i] = tmp[i - k0].second; }
                    for (uint64_t i = k0 + 1; i < k; ++i)
                        if (L->feat_bidx[i] == L->feat_bidx[i - 1]) { uint64_t exp = UINT64_MAX; dup_row.compare_exchange_strong(exp, r); }
                }
                RowStat st = {0xFFFFFFFFu, 0u, 0, 0};
                if (k > k0) {      // entries are sorted: the row's ciphertext range comes from its first and last entry
                    my_ct_min = std::min(my_ct_min, NR == 1 ? L->feat_bidx[k0] : L->feat_bidx[k0] / NR);
                    my_ct_max = std::max(my_ct_max, NR == 1 ? L->feat_bidx[k - 1] : L->feat_bidx[k - 1] / NR);
                    my_has = true;
                }
                for (uint64_t i = k0; i < k; ++i) {
                    if (NR != 1 && (((L->feat_bidx[i] % NR) * RS) & 3u)) my_unaligned = true;
                    const int32_t c = L->feat_coef[i];
                    if (c == 0) continue;
                    st.fmin = std::min(st.fmin, L->feat_bidx[i]); st.fmax = std::max(st.fmax, L->feat_bidx[i]);
                    st.cmin = std::min(st.cmin, c); st.cmax = std::max(st.cmax, c);
                }
                stat[r] = st;
            }
            wstat[w].ct_min = my_ct_min; wstat[w].ct_max = my_ct_max; wstat[w].has = my_has;
            if (my_unaligned) unaligned.store(true, std::memory_order_relaxed);
        });
        if (dup_row != UINT64_MAX) {
            const uint64_t r = dup_row;
            delete L;
            return set_error(IDASH_B200_ERR_INVALID, "layout_compile: row %llu lists an input bigIndex twice", (unsigned long long) r);
        }
        L->shifts_aligned = !unaligned;
        for (unsigned w = 0; w <= n_workers; ++w)
            if (wstat[w].has) {
                if (L->ct_min > L->ct_max) { L->ct_min = wstat[w].ct_min; L->ct_max = wstat[w].ct_max; }
                L->ct_min = std::min(L->ct_min, wstat[w].ct_min);
                L->ct_max = std::max(L->ct_max, wstat[w].ct_max);
            }

        lap("fill rows");
        // ---- band tiles for the tensor-core kernels (include/idash_b200_layout.h). Tile t = sorted rows [64 t, 64 t + 64). A row
        // whose coefficients leave the limb range, or that does not fit into the tile's band of at most RING_KMAX features, is
        // EVICTED from its tile (a hole: tile_rows = NO_ROW) and evaluated by the IMAD kernel through the overflow groups: one
        // outlier window never changes the kernel for the other rows.
        const uint32_t TN = IDASH_B200_TILE_ROWS;
        const uint64_t n_tiles_all = (n_rows + TN - 1) / TN;
        struct TilePlan { uint32_t fmin, K; uint64_t keep; uint32_t n_keep; };
        std::vector<TilePlan> plan(n_tiles_all);
        parallel_ranges(n_tiles_all, 64, [&](uint64_t t0, uint64_t t1, unsigned) {
            for (uint64_t t = t0; t < t1; ++t) {
                const uint64_t r0 = t * TN, r1 = std::min<uint64_t>(n_rows, r0 + TN);
                uint64_t keep = 0;
                uint32_t fmin = 0xFFFFFFFFu, fmax = 0;
                for (uint64_t i = r0; i < r1; ++i) {
                    const RowStat &st = stat[order[i]];
                    if (st.cmin < IDASH_B200_TILE_COEF_MIN || st.cmax > IDASH_B200_TILE_COEF_MAX) continue;
                    if (st.fmin <= st.fmax && (uint64_t) st.fmax - (st.fmin & ~31u) + 1 > IDASH_B200_RING_KMAX) continue;   // too wide by itself
                    keep |= 1ull << (i - r0);
                    if (st.fmin <= st.fmax) { fmin = std::min(fmin, st.fmin); fmax = std::max(fmax, st.fmax); }
                }
                if (fmin <= fmax && (uint64_t) fmax - (fmin & ~31u) + 1 > IDASH_B200_RING_KMAX) {
                    // the union is too wide: keep the block-aligned window of RING_KMAX features that holds the most rows
                    uint32_t best_start = fmin & ~31u, best_cnt = 0;
                    for (uint64_t i = r0; i < r1; ++i) {
                        if (!((keep >> (i - r0)) & 1u)) continue;
                        const RowStat &si = stat[order[i]];
                        if (si.fmin > si.fmax) continue;
                        const uint32_t start = si.fmin & ~31u;
                        uint32_t cnt = 0;
                        for (uint64_t j = r0; j < r1; ++j) {
                            if (!((keep >> (j - r0)) & 1u)) continue;
                            const RowStat &sj = stat[order[j]];
                            if (sj.fmin > sj.fmax || (sj.fmin >= start && (uint64_t) sj.fmax < (uint64_t) start + IDASH_B200_RING_KMAX)) ++cnt;
                        }
                        if (cnt > best_cnt) { best_cnt = cnt; best_start = start; }
                    }
                    fmin = 0xFFFFFFFFu; fmax = 0;
                    for (uint64_t i = r0; i < r1; ++i) {
                        if (!((keep >> (i - r0)) & 1u)) continue;
                        const RowStat &si = stat[order[i]];
                        if (si.fmin > si.fmax) continue;
                        if (si.fmin >= best_start && (uint64_t) si.fmax < (uint64_t) best_start + IDASH_B200_RING_KMAX) { fmin = std::min(fmin, si.fmin); fmax = std::max(fmax, si.fmax); }
                        else keep &= ~(1ull << (i - r0));
                    }
                }
                TilePlan &P = plan[t];
                P.keep = keep;
                P.n_keep = (uint32_t) __builtin_popcountll(keep);
                if (fmin > fmax) { P.fmin = 0xFFFFFFFFu; P.K = 32; }      // bias-only tile: placed by the serial pass below
                else { P.fmin = fmin & ~31u; P.K = (uint32_t) (((uint64_t) fmax - P.fmin + 1 + 31) / 32 * 32); }
            }
        });
        lap("plan tiles");
        // serial: drop empty tiles, place bias-only tiles, offsets, ring eligibility (block-aligned bands that only move forward)
        std::vector<uint64_t> tile_src;     // surviving tile -> index in plan
        L->ring_ok = true;
        uint64_t coef_bytes = 0, used_words = 0, f_end = 0;
        for (uint64_t t = 0; t < n_tiles_all; ++t) {
            TilePlan &P = plan[t];
            if (P.n_keep == 0) continue;
            if (P.fmin == 0xFFFFFFFFu) P.fmin = L->tiles.empty() ? 0u : L->tiles.back().f_base;    // bias-only: stay in place
            if ((uint64_t) P.fmin + P.K > 0xFFFFFFFFull) { P.n_keep = 0; P.keep = 0; continue; }
            if (!L->tiles.empty()) {
                const idash_b200_tile &prev = L->tiles.back();
                // a band that ends before its predecessor's is stretched to the same end when that keeps it a legal tile
                if ((uint64_t) P.fmin + P.K < (uint64_t) prev.f_base + prev.K && P.fmin >= prev.f_base &&
                    (uint64_t) prev.f_base + prev.K - P.fmin <= IDASH_B200_RING_KMAX)
                    P.K = prev.f_base + prev.K - P.fmin;
                if (P.fmin < prev.f_base || (uint64_t) P.fmin + P.K < (uint64_t) prev.f_base + prev.K) L->ring_ok = false;
            }
            idash_b200_tile T;
            memset(&T, 0, sizeof(T));
            T.f_base = P.fmin;
            T.K = P.K;
            T.b_off = coef_bytes;
            T.used_off = (uint32_t) used_words;
            T.n_valid = (uint32_t) (std::min<uint64_t>(n_rows, t * TN + TN) - t * TN);
            coef_bytes += 2ull * P.K * TN;
            used_words += P.K / 32;
            f_end = std::max<uint64_t>(f_end, (uint64_t) T.f_base + T.K);
            L->tile_kmax = std::max<uint32_t>(L->tile_kmax, T.K);
            L->tiles.push_back(T);
            tile_src.push_back(t);
        }
        if (used_words > 0xFFFFFFFFull) { delete L; return set_error(IDASH_B200_ERR_INVALID, "layout_compile: too many tiles"); }
        const uint64_t n_tiles = L->tiles.size();
        lap("place tiles");
        L->tile_rows.resize(n_tiles * TN);      // (uninitialised: every tile clears its own part in the parallel pass below)
        L->tile_bias.resize(n_tiles * TN);
        L->tile_coef.resize(coef_bytes);
        L->tile_used.resize(used_words);
        lap("alloc tiles");
        parallel_ranges(n_tiles, 32, [&](uint64_t t0, uint64_t t1, unsigned) {
            for (uint64_t ti = t0; ti < t1; ++ti) {
                idash_b200_tile &T = L->tiles[ti];
                const uint64_t src = tile_src[ti], r0 = src * TN, r1 = std::min<uint64_t>(n_rows, r0 + TN);
                const uint64_t keep = plan[src].keep;
                uint8_t *img = L->tile_coef.data() + T.b_off;
                uint32_t *used = L->tile_used.data() + T.used_off;
                memset(img, 0, 2ull * T.K * TN);
                memset(used, 0, (T.K / 32) * sizeof(uint32_t));
                for (uint32_t n = 0; n < TN; ++n) { L->tile_rows[ti * TN + n] = IDASH_B200_NO_ROW; L->tile_bias[ti * TN + n] = 0; }
                // bit 0 of flags: a full tile without holes whose caller rows are consecutive integers (fast store addressing)
                bool contiguous = (r1 - r0 == TN) && keep == ~0ull;
                for (uint64_t i = r0; i < r1 && contiguous; ++i) contiguous = order[i] == order[r0] + (i - r0);
                T.flags = contiguous ? 1u : 0u;
                for (uint64_t i = r0; i < r1; ++i) {
                    if (!((keep >> (i - r0)) & 1u)) continue;
                    const uint32_t r = order[i], n = (uint32_t) (i - r0);
                    L->tile_rows[ti * TN + n] = r;
                    L->tile_bias[ti * TN + n] = L->bias[r];
                    for (uint64_t e = L->feat_ptr[r]; e < L->feat_ptr[r + 1]; ++e) {
                        const int32_t c = L->feat_coef[e];
                        if (c == 0) continue;
                        const uint32_t k = L->feat_bidx[e] - T.f_base;
                        // chunk of 32 features = 4096 bytes = two 16-feature halves of [c_lo rows (1024 B) | c_hi rows (1024 B)]
                        const size_t o = (size_t) (k / 32) * (2 * 32 * TN) + (size_t) ((k % 32) / 16) * (2 * TN * 16) + (size_t) n * 16 + (k % 16);
                        const int32_t c_lo = ((c + 128) & 0xFF) - 128;       // balanced signed limbs: coef = c_lo + 256 c_hi
                        const int32_t c_hi = (c - c_lo) >> 8;
                        img[o] = (uint8_t) (int8_t) c_lo;
                        img[o + TN * 16] = (uint8_t) (int8_t) c_hi;
                        used[k / 32] |= 1u << (k % 32);
                    }
                }
            }
        });
        lap("fill tiles");
        if (f_end > (1ull << 30)) L->ring_ok = false;
        if (n_tiles == 0) L->ring_ok = false;
        if (L->ring_ok) {
            L->feat_used.assign(f_end / 32, 0);
            for (const idash_b200_tile &T : L->tiles)
                for (uint32_t w = 0; w < T.K / 32; ++w) L->feat_used[T.f_base / 32 + w] |= L->tile_used[T.used_off + w];
        }

        // ---- overflow rows -> IMAD groups (or every row when the caller wants the IMAD layout of the whole model)
        std::vector<uint32_t> overflow;
        for (uint64_t t = 0; t < n_tiles_all; ++t) {
            const uint64_t r0 = t * TN, r1 = std::min<uint64_t>(n_rows, r0 + TN);
            if (plan[t].n_keep == r1 - r0) continue;
            for (uint64_t i = r0; i < r1; ++i)
                if (!((plan[t].keep >> (i - r0)) & 1u)) overflow.push_back(order[i]);
        }
        L->n_overflow_rows = overflow.size();
        if (!overflow.empty()) build_groups(L, overflow, L->groups, L->entries);
        if (flags & IDASH_B200_COMPILE_GROUPS_ALL) {
            build_groups(L, L->order, L->groups_full, L->entries_full);
            L->groups_all = true;
        }
        lap("groups");
    } catch (const std::bad_alloc &) {
        delete L;
        return set_error(IDASH_B200_ERR_NOMEM, "layout_compile: out of memory");
    }
    *out = L;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_compile(const idash_b200_model_desc *d, idash_b200_layout **out) {
    return idash_b200_layout_compile_ex(d, IDASH_B200_COMPILE_GROUPS_ALL, out);
}

extern "C" int idash_b200_layout_ensure_groups_all(idash_b200_layout *L) {
    clear_error();
    if (!L) return set_error(IDASH_B200_ERR_INVALID, "layout_ensure_groups_all: null argument");
    if (L->groups_all) return IDASH_B200_OK;
    try {
        build_groups(L, L->order, L->groups_full, L->entries_full);
    } catch (const std::bad_alloc &) {
        return set_error(IDASH_B200_ERR_NOMEM, "layout_ensure_groups_all: out of memory");
    }
    L->groups_all = true;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_free(idash_b200_layout *layout) {
    if (layout && layout->refs.fetch_sub(1) == 1) delete layout;
    return IDASH_B200_OK;
}

// ---- cached packed model ("models.bin"): the compiled layout as one file ------------------------------------------
namespace {
const uint64_t kMagic = 0x314C444D30303242ull;   // "B200MDL1"
const uint32_t kVersion = 3;

struct FileHeader {
    uint64_t magic;
    uint32_t version, tile_rows;
    uint64_t key;
    uint32_t S, NR, RS, tile_kmax;
    uint64_t n_rows, nnz, n_overflow_rows;
    uint32_t ct_min, ct_max, max_entries_per_group;
    uint8_t ring_ok, groups_all, shifts_aligned, pad;
    uint64_t n[20];    // element counts of the arrays, in the order of for_each_array
};

template <class F>
void for_each_array(idash_b200_layout *L, F f) {
    f(0, L->out_bidx); f(1, L->feat_ptr); f(2, L->feat_bidx); f(3, L->feat_coef); f(4, L->bias); f(5, L->order);
    f(6, L->var_ptr); f(7, L->var_ct); f(8, L->var_w); f(9, L->var_wsum); f(10, L->tiles); f(11, L->tile_rows);
    f(12, L->tile_bias); f(13, L->tile_coef); f(14, L->tile_used); f(15, L->feat_used); f(16, L->groups); f(17, L->entries);
    f(18, L->groups_full); f(19, L->entries_full);
}
}  // namespace

extern "C" int idash_b200_layout_save(const idash_b200_layout *Lc, const char *path, uint64_t key) {
    clear_error();
    if (!Lc || !path) return set_error(IDASH_B200_ERR_INVALID, "layout_save: null argument");
    idash_b200_layout *L = const_cast<idash_b200_layout *>(Lc);
    FileHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kMagic; h.version = kVersion; h.tile_rows = IDASH_B200_TILE_ROWS; h.key = key;
    h.S = L->S; h.NR = L->NR; h.RS = L->RS; h.tile_kmax = L->tile_kmax;
    h.n_rows = L->n_rows; h.nnz = L->nnz; h.n_overflow_rows = L->n_overflow_rows;
    h.ct_min = L->ct_min; h.ct_max = L->ct_max; h.max_entries_per_group = L->max_entries_per_group;
    h.ring_ok = L->ring_ok; h.groups_all = L->groups_all; h.shifts_aligned = L->shifts_aligned;
    for_each_array(L, [&](int i, auto &v) { h.n[i] = v.size(); });
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return set_error(IDASH_B200_ERR_INVALID, "layout_save: cannot open %s for write", tmp.c_str());
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    for_each_array(L, [&](int, auto &v) {
        if (ok && !v.empty()) ok = fwrite(v.data(), sizeof(v[0]), v.size(), f) == v.size();
    });
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return set_error(IDASH_B200_ERR_INVALID, "layout_save: short write to %s", path); }
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_load(const char *path, uint64_t key, idash_b200_layout **out) {
    clear_error();
    if (!path || !out) return set_error(IDASH_B200_ERR_INVALID, "layout_load: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return set_error(IDASH_B200_ERR_INVALID, "layout_load: cannot open %s", path);
    FileHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || h.magic != kMagic || h.version != kVersion || h.tile_rows != IDASH_B200_TILE_ROWS || h.key != key) {
        fclose(f);
        return set_error(IDASH_B200_ERR_INVALID, "layout_load: %s is not a cached model for this key / library version", path);
    }
    idash_b200_layout *L = new (std::nothrow) idash_b200_layout();
    if (!L) { fclose(f); return set_error(IDASH_B200_ERR_NOMEM, "layout_load: out of memory"); }
    bool ok = true;
    try {
        L->S = h.S; L->NR = h.NR; L->RS = h.RS; L->tile_kmax = h.tile_kmax;
        L->n_rows = h.n_rows; L->nnz = h.nnz; L->n_overflow_rows = h.n_overflow_rows;
        L->ct_min = h.ct_min; L->ct_max = h.ct_max; L->max_entries_per_group = h.max_entries_per_group;
        L->ring_ok = h.ring_ok != 0; L->groups_all = h.groups_all != 0; L->shifts_aligned = h.shifts_aligned != 0;
        // sizes are checked against the file length before anything is allocated
        fseek(f, 0, SEEK_END);
        const uint64_t file_bytes = (uint64_t) ftell(f);
        fseek(f, (long) sizeof(h), SEEK_SET);
        uint64_t need = sizeof(h);
        for_each_array(L, [&](int i, auto &v) { need += h.n[i] * sizeof(v[0]); });
        ok = need == file_bytes && h.n[0] == h.n_rows && h.n[1] == h.n_rows + 1 && h.n[6] == h.n_rows + 1 && h.n[4] == h.n_rows &&
             h.n[5] == h.n_rows && h.n[9] == h.n_rows && h.n[11] == h.n[10] * IDASH_B200_TILE_ROWS && h.n[12] == h.n[11];
        if (ok)
            for_each_array(L, [&](int i, auto &v) {
                v.resize(h.n[i]);
                if (ok && h.n[i]) ok = fread(v.data(), sizeof(v[0]), h.n[i], f) == h.n[i];
            });
    } catch (const std::bad_alloc &) {
        ok = false;
    }
    fclose(f);
    if (!ok) { delete L; return set_error(IDASH_B200_ERR_INVALID, "layout_load: %s is truncated or inconsistent", path); }
    *out = L;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_get_info(const idash_b200_layout *L, idash_b200_model_info *info) {
    if (!L || !info) return set_error(IDASH_B200_ERR_INVALID, "layout_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->n_rows = L->n_rows;
    info->nnz = L->nnz;
    info->n_groups = L->groups_all ? L->groups_full.size() : L->groups.size();
    info->n_entries = L->groups_all ? L->entries_full.size() : L->entries.size();
    info->ct_min = L->ct_min;
    info->ct_max = L->ct_max;
    info->max_entries_per_group = L->max_entries_per_group;
    info->shifts_aligned = L->shifts_aligned ? 1u : 0u;
    info->n_tiles = L->tiles.size();
    info->ring_ok = L->ring_ok ? 1u : 0u;
    info->tile_kmax = L->tile_kmax;
    info->n_overflow_rows = L->n_overflow_rows;
    info->groups_all = L->groups_all ? 1u : 0u;
    info->device_bytes = L->tiles.size() * sizeof(idash_b200_tile) + L->tile_rows.size() * 4 + L->tile_bias.size() * 4 +
                         L->tile_coef.size() + L->tile_used.size() * 4 + (L->groups.size() + L->groups_full.size()) * sizeof(idash_b200_group) + (L->entries.size() + L->entries_full.size()) * sizeof(idash_b200_entry) +
                         L->var_ptr.size() * 8 + L->var_ct.size() * 4 + L->var_w.size() * 8 + L->var_wsum.size() * 8 + L->out_bidx.size() * 4;
    return IDASH_B200_OK;
}

extern "C" const idash_b200_group *idash_b200_layout_groups(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->groups_full.size();
    return L->groups_full.data();
}
extern "C" const idash_b200_entry *idash_b200_layout_entries(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->entries_full.size();
    return L->entries_full.data();
}
extern "C" const idash_b200_group *idash_b200_layout_overflow_groups(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->groups.size();
    return L->groups.data();
}
extern "C" const idash_b200_entry *idash_b200_layout_overflow_entries(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->entries.size();
    return L->entries.data();
}
extern "C" const uint64_t *idash_b200_layout_feat_ptr(const idash_b200_layout *L) { return L->feat_ptr.data(); }
extern "C" const uint32_t *idash_b200_layout_feat_bidx(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->feat_bidx.size();
    return L->feat_bidx.data();
}
extern "C" const int32_t *idash_b200_layout_feat_coef(const idash_b200_layout *L) { return L->feat_coef.data(); }
extern "C" const int32_t *idash_b200_layout_bias(const idash_b200_layout *L) { return L->bias.data(); }
extern "C" const uint64_t *idash_b200_layout_var_ptr(const idash_b200_layout *L) { return L->var_ptr.data(); }
extern "C" const uint32_t *idash_b200_layout_var_ct(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->var_ct.size();
    return L->var_ct.data();
}
extern "C" const double *idash_b200_layout_var_w(const idash_b200_layout *L) { return L->var_w.data(); }
extern "C" const uint32_t *idash_b200_layout_out_bidx(const idash_b200_layout *L) { return L->out_bidx.data(); }
extern "C" const idash_b200_tile *idash_b200_layout_tiles(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->tiles.size();
    return L->tiles.data();
}
extern "C" const uint32_t *idash_b200_layout_tile_rows(const idash_b200_layout *L) { return L->tile_rows.data(); }
extern "C" const int32_t *idash_b200_layout_tile_bias(const idash_b200_layout *L) { return L->tile_bias.data(); }
extern "C" const uint8_t *idash_b200_layout_tile_coef(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->tile_coef.size();
    return L->tile_coef.data();
}
extern "C" const uint32_t *idash_b200_layout_tile_used(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->tile_used.size();
    return L->tile_used.data();
}
extern "C" const uint32_t *idash_b200_layout_feat_used(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->ring_ok ? L->feat_used.size() : 0;
    return L->feat_used.data();
}
