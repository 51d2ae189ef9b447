// Model -> device block-banded layout compiler (host, pure C++; no CUDA in this file).
//
// Replaces, for the B200 path, what the reference does on every call of cloud_compute_score:
// deep-copying Model::model into a vector (eval/idash.cpp:772) and walking per-output hash maps of
// (input bigIndex -> coefficient) (eval/idash.cpp:800-819). See include/idash_b200_layout.h for
// the layout and DESIGN.md for why it has this shape.
#include <algorithm>
#include <cstring>
#include <new>
#include <numeric>

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

namespace {

struct Feat {          // one non-constant model entry of one row
    uint32_t bidx;     // input bigIndex
    int32_t coef;
};

struct Triple {        // the (up to) three variant rows of one target SNP
    uint32_t target;   // out_bidx / 3
    int64_t row[3] = {-1, -1, -1};      // caller row per variant
};

}  // namespace
}  // namespace idash_b200

using namespace idash_b200;

extern "C" const char *idash_b200_last_error(void) { return idash_b200::g_err; }

extern "C" int idash_b200_layout_compile(const idash_b200_model_desc *d, idash_b200_layout **out) {
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

    idash_b200_layout *L = new (std::nothrow) idash_b200_layout();
    if (!L) return set_error(IDASH_B200_ERR_NOMEM, "layout_compile: out of memory");
    try {
        L->S = S; L->NR = NR; L->RS = RS; L->n_rows = n_rows; L->nnz = nnz;
        L->out_bidx.assign(d->out_bidx, d->out_bidx + n_rows);

        // rows sorted by output bigIndex (genomic order when the targets file is sorted)
        std::vector<uint32_t> order(n_rows);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return d->out_bidx[a] < d->out_bidx[b]; });
        for (uint64_t i = 1; i < n_rows; ++i)
            if (d->out_bidx[order[i]] == d->out_bidx[order[i - 1]]) {
                delete L;
                return set_error(IDASH_B200_ERR_INVALID, "layout_compile: duplicate output bigIndex %u", d->out_bidx[order[i]]);
            }

        // per-row: bias, sorted features, variance terms
        std::vector<int32_t> bias(n_rows, 0);
        std::vector<std::vector<Feat>> feats(n_rows);
        L->var_ptr.assign(n_rows + 1, 0);
        for (uint64_t r = 0; r < n_rows; ++r) {
            auto &fv = feats[r];
            bool have_const = false;
            for (uint64_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e) {
                if (d->col[e] == IDASH_B200_CONSTANT_BIDX) {
                    if (have_const) { delete L; return set_error(IDASH_B200_ERR_INVALID, "layout_compile: row %llu has two Constant entries", (unsigned long long) r); }
                    have_const = true;
                    bias[r] = d->coef[e];
                } else {
                    fv.push_back({d->col[e], d->coef[e]});
                }
            }
            std::sort(fv.begin(), fv.end(), [](const Feat &a, const Feat &b) { return a.bidx < b.bidx; });
            for (size_t i = 1; i < fv.size(); ++i)
                if (fv[i].bidx == fv[i - 1].bidx) { delete L; return set_error(IDASH_B200_ERR_INVALID, "layout_compile: row %llu lists input bigIndex %u twice", (unsigned long long) r, fv[i].bidx); }
            for (const Feat &f : fv) {
                const uint32_t ct = f.bidx / NR, region = f.bidx % NR;
                if (L->ct_min > L->ct_max) { L->ct_min = L->ct_max = ct; }
                L->ct_min = std::min(L->ct_min, ct);
                L->ct_max = std::max(L->ct_max, ct);
                if (region == 0) {
                    // tLweAddMulTo: current_variance += (p * p) * sample->current_variance with an int32 p * p
                    const int32_t pp = (int32_t) ((uint32_t) f.coef * (uint32_t) f.coef);
                    L->var_ct.push_back(ct);
                    L->var_w.push_back((double) pp);
                }
            }
            L->var_ptr[r + 1] = L->var_ct.size();
        }

        // triples: rows of the same target SNP (out_bidx / 3)
        std::vector<Triple> triples;
        for (uint64_t i = 0; i < n_rows; ++i) {
            const uint32_t r = order[i];
            const uint32_t target = d->out_bidx[r] / 3, variant = d->out_bidx[r] % 3;
            if (triples.empty() || triples.back().target != target) {
                Triple t;
                t.target = target;
                triples.push_back(t);
            }
            triples.back().row[variant] = r;
        }

        // groups of two triples; entries = union of (ct, shift), A-only | shared | B-only
        const uint64_t n_groups = (triples.size() + 1) / 2;
        L->groups.resize(n_groups);
        struct Acc { uint32_t bidx; int32_t c[6]; uint8_t used; };
        std::vector<Acc> uni;
        for (uint64_t g = 0; g < n_groups; ++g) {
            idash_b200_group &G = L->groups[g];
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
                    G.bias[slot] = bias[r];
                    for (const Feat &f : feats[r]) {
                        if (f.coef == 0) continue;   // contributes nothing to words or variance
                        Acc a;
                        a.bidx = f.bidx;
                        memset(a.c, 0, sizeof(a.c));
                        a.c[slot] = f.coef;
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
            if (L->entries.size() + m > 0xFFFFFFFFull) { delete L; return set_error(IDASH_B200_ERR_INVALID, "layout_compile: too many entries"); }
            G.entry_begin = (uint32_t) L->entries.size();
            for (int cls = 1; cls <= 3; ++cls) {          // used == 1: A only, 3: both, 2: B only
                const uint8_t want = cls == 1 ? 1 : (cls == 2 ? 3 : 2);
                uint32_t cnt = 0;
                for (const Acc &a : uni) {
                    if (a.used != want) continue;
                    idash_b200_entry E;
                    E.ct = a.bidx / NR;
                    E.shift = (a.bidx % NR) * RS;
                    if (E.shift & 3u) L->shifts_aligned = false;
                    memcpy(E.coef, a.c, sizeof(E.coef));
                    L->entries.push_back(E);
                    ++cnt;
                }
                if (cls == 1) G.n_a = cnt; else if (cls == 2) G.n_ab = cnt; else G.n_b = cnt;
            }
            L->max_entries_per_group = std::max<uint32_t>(L->max_entries_per_group, (uint32_t) m);
        }

        // ---- band tiles for the tensor-core kernel (include/idash_b200_layout.h) ----
        {
            const uint32_t TN = IDASH_B200_TILE_ROWS;
            const uint64_t n_tiles = (n_rows + TN - 1) / TN;
            bool ok = true;
            L->tiles.resize(n_tiles);
            L->tile_rows.assign(n_tiles * TN, IDASH_B200_NO_ROW);
            L->tile_bias.assign(n_tiles * TN, 0);
            for (uint64_t t = 0; t < n_tiles && ok; ++t) {
                const uint64_t r0 = t * TN, r1 = std::min<uint64_t>(n_rows, r0 + TN);
                uint32_t fmin = 0xFFFFFFFFu, fmax = 0;
                for (uint64_t i = r0; i < r1 && ok; ++i)
                    for (const Feat &f : feats[order[i]]) {
                        if (f.coef == 0) continue;
                        if (f.coef < IDASH_B200_TILE_COEF_MIN || f.coef > IDASH_B200_TILE_COEF_MAX) { ok = false; break; }
                        fmin = std::min(fmin, f.bidx);
                        fmax = std::max(fmax, f.bidx);
                    }
                if (!ok) break;
                if (fmin > fmax) fmin = fmax = (t ? L->tiles[t - 1].f_base : 0);   // bias-only tile: stay in place
                // bands start on a 32-feature block so that consecutive tiles share whole staged blocks (persistent ring kernel)
                fmin &= ~31u;
                const uint64_t width = (uint64_t) fmax - fmin + 1;
                const uint64_t K = (width + 31) / 32 * 32;
                if (K > IDASH_B200_TILE_KMAX || (uint64_t) fmin + K > 0xFFFFFFFFull) { ok = false; break; }
                idash_b200_tile &T = L->tiles[t];
                memset(&T, 0, sizeof(T));
                T.f_base = fmin;
                T.K = (uint32_t) K;
                T.b_off = L->tile_coef.size();
                T.used_off = (uint32_t) L->tile_used.size();
                T.n_valid = (uint32_t) (r1 - r0);
                {   // bit 0 of flags: a full tile whose caller rows are consecutive integers (fast store addressing)
                    bool contiguous = (r1 - r0 == TN);
                    for (uint64_t i = r0; i < r1 && contiguous; ++i) contiguous = order[i] == order[r0] + (i - r0);
                    T.flags = contiguous ? 1u : 0u;
                }
                L->tile_kmax = std::max<uint32_t>(L->tile_kmax, T.K);
                L->tile_coef.resize(L->tile_coef.size() + 2 * K * TN, 0);
                L->tile_used.resize(L->tile_used.size() + K / 32, 0);
                uint8_t *img = L->tile_coef.data() + T.b_off;
                uint32_t *used = L->tile_used.data() + T.used_off;
                for (uint64_t i = r0; i < r1; ++i) {
                    const uint32_t r = order[i], n = (uint32_t) (i - r0);
                    L->tile_rows[t * TN + n] = r;
                    L->tile_bias[t * TN + n] = bias[r];
                    for (const Feat &f : feats[r]) {
                        if (f.coef == 0) continue;
                        const uint32_t k = f.bidx - fmin;
                        // chunk of 32 features = 4096 bytes = two 16-feature halves of [c_lo rows (1024 B) | c_hi rows (1024 B)]
                        const size_t o = (size_t) (k / 32) * (2 * 32 * TN) + (size_t) ((k % 32) / 16) * (2 * TN * 16) + (size_t) n * 16 + (k % 16);
                        const int32_t c_lo = ((f.coef + 128) & 0xFF) - 128;       // balanced signed limbs: coef = c_lo + 256 c_hi
                        const int32_t c_hi = (f.coef - c_lo) >> 8;
                        img[o] = (uint8_t) (int8_t) c_lo;
                        img[o + TN * 16] = (uint8_t) (int8_t) c_hi;
                        used[k / 32] |= 1u << (k % 32);
                    }
                }
            }
            if (!ok) {
                L->tiles.clear(); L->tile_rows.clear(); L->tile_bias.clear(); L->tile_coef.clear(); L->tile_used.clear();
                L->tile_kmax = 0;
            }
            // persistent ring kernel: block-aligned bands that only move forward, at most
            // IDASH_B200_RING_KMAX features wide; feat_used = features some row multiplies by a non-zero coefficient
            L->ring_ok = ok && !L->tiles.empty() && L->tile_kmax <= IDASH_B200_RING_KMAX;
            uint64_t f_end = 0;
            for (uint64_t t = 0; t < L->tiles.size() && L->ring_ok; ++t) {
                const idash_b200_tile &T = L->tiles[t];
                if (t && (T.f_base < L->tiles[t - 1].f_base || T.f_base + T.K < L->tiles[t - 1].f_base + L->tiles[t - 1].K))
                    L->ring_ok = false;
                f_end = std::max<uint64_t>(f_end, (uint64_t) T.f_base + T.K);
            }
            if (f_end > (1ull << 30)) L->ring_ok = false;
            if (L->ring_ok) {
                L->feat_used.assign(f_end / 32, 0);
                for (const idash_b200_tile &T : L->tiles)
                    for (uint32_t w = 0; w < T.K / 32; ++w) L->feat_used[T.f_base / 32 + w] |= L->tile_used[T.used_off + w];
            }
        }
    } catch (const std::bad_alloc &) {
        delete L;
        return set_error(IDASH_B200_ERR_NOMEM, "layout_compile: out of memory");
    }
    *out = L;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_free(idash_b200_layout *layout) {
    delete layout;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_layout_get_info(const idash_b200_layout *L, idash_b200_model_info *info) {
    if (!L || !info) return set_error(IDASH_B200_ERR_INVALID, "layout_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->n_rows = L->n_rows;
    info->nnz = L->nnz;
    info->n_groups = L->groups.size();
    info->n_entries = L->entries.size();
    info->ct_min = L->ct_min;
    info->ct_max = L->ct_max;
    info->max_entries_per_group = L->max_entries_per_group;
    info->shifts_aligned = L->shifts_aligned ? 1u : 0u;
    info->n_tiles = L->tiles.size();
    info->ring_ok = L->ring_ok ? 1u : 0u;
    info->tile_kmax = L->tile_kmax;
    info->device_bytes = L->tiles.size() * sizeof(idash_b200_tile) + L->tile_rows.size() * 4 + L->tile_bias.size() * 4 +
                         L->tile_coef.size() + L->tile_used.size() * 4 + L->groups.size() * sizeof(idash_b200_group) + L->entries.size() * sizeof(idash_b200_entry) +
                         L->var_ptr.size() * 8 + L->var_ct.size() * 4 + L->var_w.size() * 8 + L->out_bidx.size() * 4;
    return IDASH_B200_OK;
}

extern "C" const idash_b200_group *idash_b200_layout_groups(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->groups.size();
    return L->groups.data();
}
extern "C" const idash_b200_entry *idash_b200_layout_entries(const idash_b200_layout *L, uint64_t *n) {
    if (n) *n = L->entries.size();
    return L->entries.data();
}
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
