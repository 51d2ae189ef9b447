// libidash_b200.so -- CUDA (sm_100a) implementation of the C ABI in include/idash_b200.h.
//
// Kernels
//   cloud_eval_kernel   K1: out = bias + sum coef * rot(in) mod 2^32 over the block-banded layout
//                           (restates eval/idash.cpp:763-848 + tlwe-functions.cpp:106-176 +
//                           toruspolynomial-functions.cpp:97-103,140-160 as one fused pass)
//   cloud_finalize_kernel   per-row variance (tlwe-functions.cpp:175) and record headers
//   slot_map_kernel         ciphertext index -> input slot table (EncryptedData::enc_data lookup,
//                           eval/idash.h:162-166)
//   decrypt_kernel      K4: phase = b - key*a (exact negacyclic, tlwe-functions.cpp:64-71 with the
//                           integer product of multiplication.cpp:53-65) fused with the decode of
//                           eval/idash.cpp:717-719
// There is no CPU fallback: every entry point fails with IDASH_B200_ERR_CUDA when no device works.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "internal.h"

using namespace idash_b200;

#define CT_WORDS 2048u
#define POLY_N 1024u
#define NO_SLOT 0xFFFFFFFFu

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return set_error(IDASH_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side views
// ------------------------------------------------------------------------------------------------
struct CtView {              // where ciphertext slot s keeps its 2048 words / variance / header
    uint8_t *words;          // address of word 0 of slot 0 (16-byte aligned)
    uint64_t stride;         // bytes between slots (8192 packed, 8208 records)
    uint32_t *index;         // PACKED: index array or null. RECORDS: null (header is words - 8)
    double *variance;        // PACKED: variance array or null. RECORDS: null (trailer is words + 8192)
    uint32_t records;        // 1 = RECORDS layout
    uint64_t count;
};

struct CloudParams {
    const idash_b200_group *groups;
    const idash_b200_entry *entries;
    uint32_t n_groups;
    uint32_t groups_per_cta;
    CtView in, out;
    const uint32_t *slot_of_ct;   // null: identity
    uint32_t n_ct_slots;          // size of slot_of_ct (or in.count for identity)
    const uint32_t *slot_of_row;  // null: identity
    uint32_t S, RS;
    int *status;
};

__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// streaming 128-bit store: outputs are written once and never re-read by this kernel
__device__ __forceinline__ void stg128_stream(void *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t lookup_slot(const CloudParams &p, uint32_t ct) {
    if (ct >= p.n_ct_slots) return NO_SLOT;
    return p.slot_of_ct ? __ldg(p.slot_of_ct + ct) : ct;
}

// MODE 0: no rotation anywhere in the model (NUM_REGIONS == 1)
// MODE 1: rotations are multiples of 4 words  -> one 128-bit load, one sign for the 4 words
// MODE 2: arbitrary rotations (e.g. REGION_SIZE = 341) -> four 32-bit loads
// Returns X^(-shift) * P restricted to words i0..i0+3 of the polynomial starting at `poly`
// (torusPolynomialMulByXai with a = 2N - shift, toruspolynomial-functions.cpp:140-160).
template <int MODE>
__device__ __forceinline__ uint4 load_rotated(const uint8_t *poly, uint32_t i0, uint32_t shift) {
    if (MODE == 0) {
        return ldg128(poly + 4u * i0);
    } else if (MODE == 1) {
        uint32_t idx = i0 + shift;
        const bool neg = idx >= POLY_N;
        idx &= POLY_N - 1;
        uint4 v = ldg128(poly + 4u * idx);
        if (neg) { v.x = 0u - v.x; v.y = 0u - v.y; v.z = 0u - v.z; v.w = 0u - v.w; }
        return v;
    } else {
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t idx = i0 + k + shift;
            const bool neg = idx >= POLY_N;
            idx &= POLY_N - 1;
            const uint32_t x = __ldg(reinterpret_cast<const uint32_t *>(poly) + idx);
            r[k] = neg ? 0u - x : x;
        }
        return make_uint4(r[0], r[1], r[2], r[3]);
    }
}

#define MAC4(acc, c, x)                      \
    do {                                     \
        (acc).x += (uint32_t) (c) * (x).x;   \
        (acc).y += (uint32_t) (c) * (x).y;   \
        (acc).z += (uint32_t) (c) * (x).z;   \
        (acc).w += (uint32_t) (c) * (x).w;   \
    } while (0)

// One CTA = 128 threads = one 512-word slice of the 2048-word ciphertext axis, walking
// `groups_per_cta` consecutive groups (consecutive target SNPs, whose tag windows overlap, so the
// re-read input words hit L1/L2). Each thread owns 4 consecutive words (128-bit accesses) and the
// 6 x 4 int32 accumulators of the group's rows.
template <int MODE>
__global__ void __launch_bounds__(128) cloud_eval_kernel(const CloudParams p) {
    const uint32_t slice = blockIdx.x & 3u;
    const uint32_t chunk = blockIdx.x >> 2;
    const uint32_t w0 = slice * 512u + threadIdx.x * 4u;   // first word of this thread inside a ct
    const uint32_t poly_off = w0 & POLY_N;                  // 0: polynomial a, 1024: polynomial b
    const uint32_t i0 = w0 & (POLY_N - 1);
    const bool is_b = poly_off != 0;

    uint32_t g = chunk * p.groups_per_cta;
    const uint32_t g_end = min(p.n_groups, g + p.groups_per_cta);
    for (; g < g_end; ++g) {
        const uint4 *gp = reinterpret_cast<const uint4 *>(p.groups + g);
        const uint4 g0 = __ldg(gp);   // entry_begin, n_a, n_ab, n_b
        uint4 acc[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r] = make_uint4(0, 0, 0, 0);

        const idash_b200_entry *ep = p.entries + g0.x;
        // ---- entries used only by target A: rows 0..2
#pragma unroll 2
        for (uint32_t k = 0; k < g0.y; ++k, ++ep) {
            const uint4 h = ldg128(ep);                                       // ct, shift, c0, c1
            const uint4 t = ldg128(reinterpret_cast<const uint4 *>(ep) + 1);  // c2, c3, c4, c5
            const uint32_t slot = lookup_slot(p, h.x);
            if (slot == NO_SLOT) { if (threadIdx.x == 0) atomicOr(p.status, 1); continue; }
            const uint4 x = load_rotated<MODE>(p.in.words + (uint64_t) slot * p.in.stride + 4u * poly_off, i0, h.y);
            MAC4(acc[0], h.z, x); MAC4(acc[1], h.w, x); MAC4(acc[2], t.x, x);
        }
        // ---- entries shared by A and B: rows 0..5
#pragma unroll 2
        for (uint32_t k = 0; k < g0.z; ++k, ++ep) {
            const uint4 h = ldg128(ep);
            const uint4 t = ldg128(reinterpret_cast<const uint4 *>(ep) + 1);
            const uint32_t slot = lookup_slot(p, h.x);
            if (slot == NO_SLOT) { if (threadIdx.x == 0) atomicOr(p.status, 1); continue; }
            const uint4 x = load_rotated<MODE>(p.in.words + (uint64_t) slot * p.in.stride + 4u * poly_off, i0, h.y);
            MAC4(acc[0], h.z, x); MAC4(acc[1], h.w, x); MAC4(acc[2], t.x, x);
            MAC4(acc[3], t.y, x); MAC4(acc[4], t.z, x); MAC4(acc[5], t.w, x);
        }
        // ---- entries used only by target B: rows 3..5
#pragma unroll 2
        for (uint32_t k = 0; k < g0.w; ++k, ++ep) {
            const uint4 h = ldg128(ep);
            const uint4 t = ldg128(reinterpret_cast<const uint4 *>(ep) + 1);
            const uint32_t slot = lookup_slot(p, h.x);
            if (slot == NO_SLOT) { if (threadIdx.x == 0) atomicOr(p.status, 1); continue; }
            const uint4 x = load_rotated<MODE>(p.in.words + (uint64_t) slot * p.in.stride + 4u * poly_off, i0, h.y);
            MAC4(acc[3], t.y, x); MAC4(acc[4], t.z, x); MAC4(acc[5], t.w, x);
        }

        // ---- epilogue: bias on b[0..S) (idash.cpp:805-810), b[RS..N) = 0 (idash.cpp:839-841), store
        const uint4 rows03 = __ldg(gp + 1);               // row[0..3]
        const uint4 rows45_b01 = __ldg(gp + 2);           // row[4], row[5], bias[0], bias[1]
        const uint4 b25 = __ldg(gp + 3);                  // bias[2..5]
        const uint32_t rows[6] = {rows03.x, rows03.y, rows03.z, rows03.w, rows45_b01.x, rows45_b01.y};
        const uint32_t bias[6] = {rows45_b01.z, rows45_b01.w, b25.x, b25.y, b25.z, b25.w};
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            if (rows[r] == IDASH_B200_NO_ROW) continue;
            uint4 v = acc[r];
            if (is_b) {
                const uint32_t sb = bias[r] * (uint32_t) IDASH_B200_ONE_IN_T32;
                v.x = (i0 + 0 >= p.RS) ? 0u : v.x + ((i0 + 0 < p.S) ? sb : 0u);
                v.y = (i0 + 1 >= p.RS) ? 0u : v.y + ((i0 + 1 < p.S) ? sb : 0u);
                v.z = (i0 + 2 >= p.RS) ? 0u : v.z + ((i0 + 2 < p.S) ? sb : 0u);
                v.w = (i0 + 3 >= p.RS) ? 0u : v.w + ((i0 + 3 < p.S) ? sb : 0u);
            }
            const uint32_t oslot = p.slot_of_row ? __ldg(p.slot_of_row + rows[r]) : rows[r];
            stg128_stream(p.out.words + (uint64_t) oslot * p.out.stride + 4u * w0, v);
        }
    }
}

#include "cloud_tc.cuh"
#include "cloud_ring.cuh"

// The production library carries one epilogue schedule; the profiling build all of them (IDASH_B200_TUNE bits 256 / 512).
#ifdef IDASH_B200_PROFILE
#define RG_EPI_VARIANTS 4
#else
#define RG_EPI_VARIANTS 1
#endif
typedef void (*ring_kernel_t)(const RingParams);
template <int EPI>
static ring_kernel_t ring_kernel_pick(bool rot, bool batched) {
    return rot ? (batched ? cloud_ring_kernel<true, true, EPI> : cloud_ring_kernel<true, false, EPI>)
               : (batched ? cloud_ring_kernel<false, true, EPI> : cloud_ring_kernel<false, false, EPI>);
}
static ring_kernel_t ring_kernel_fn(bool rot, bool batched, int epi) {
#ifdef IDASH_B200_PROFILE
    if (epi == 3) return ring_kernel_pick<3>(rot, batched);
    if (epi == 2) return ring_kernel_pick<2>(rot, batched);
    if (epi == 1) return ring_kernel_pick<1>(rot, batched);
    return ring_kernel_pick<0>(rot, batched);
#else
    (void) epi;
    return ring_kernel_pick<RG_EPI_DEFAULT>(rot, batched);
#endif
}

// Per caller row: output variance, record header / index array.
struct FinalizeParams {
    uint64_t row_lo;     // rows [row_lo, n_rows) of the model
    uint64_t n_rows;
    const uint64_t *var_ptr;
    const uint32_t *var_ct;
    const double *var_w;
    const uint32_t *out_bidx;
    CtView in, out;
    const uint32_t *slot_of_ct;
    uint32_t n_ct_slots;
    const uint32_t *slot_of_row;
    double default_var;
    const double *var_wsum;          // per row: sum of the weights (coefficient^2 of the region-0 entries)
    const uint64_t *var_uniform;     // device: [0] != 0 -> the input variances are not all equal; [1] = bits of the common value;
                                     // null -> every input has default_var
};

// Do all input ciphertexts carry the same variance? (They do when the inputs come from the reference's encrypt: alpha^2 = 2^-50
// for every ciphertext, eval/idash.cpp:625.) flag[0] is cleared by the launcher; flag[1] receives the bits of slot 0's variance.
__global__ void variance_scan_kernel(CtView in, uint64_t *flag) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.count) return;
    const unsigned long long v0 = in.records ? *reinterpret_cast<const unsigned long long *>(in.words + 8192)
                                             : (unsigned long long) __double_as_longlong(in.variance[0]);
    const unsigned long long v = in.records ? *reinterpret_cast<const unsigned long long *>(in.words + i * in.stride + 8192)
                                            : (unsigned long long) __double_as_longlong(in.variance[i]);
    if (i == 0) flag[1] = v0;
    if (v != v0) flag[0] = 1;
}

__global__ void cloud_finalize_kernel(const FinalizeParams p) {
    const uint64_t r = p.row_lo + (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_rows) return;
    double var = 0.;
    // Fast path: one common input variance that is a power of two (or zero). Then sum_e w_e * v = (sum_e w_e) * v EXACTLY (the
    // weights are integers below 2^53 in total, the scaling is exact), so the per-entry walk -- 151 entries per row at
    // neighbors = 50, read with a row-length stride -- collapses to one multiply. Anything else takes the general loop.
    bool fast = false;
    double v_common = p.default_var;
    if (p.var_uniform) {
        const unsigned long long bits = p.var_uniform[1];
        fast = p.var_uniform[0] == 0 && (bits & 0x000FFFFFFFFFFFFFull) == 0 && ((bits >> 52) & 0x7FFu) != 0x7FFu;
        v_common = __longlong_as_double((long long) bits);
    } else {
        fast = true;
    }
    if (fast) var = p.var_wsum[r] * v_common;
    else
    for (uint64_t e = p.var_ptr[r]; e < p.var_ptr[r + 1]; ++e) {
        const uint32_t ct = p.var_ct[e];
        uint32_t slot = NO_SLOT;
        if (ct < p.n_ct_slots) slot = p.slot_of_ct ? p.slot_of_ct[ct] : ct;
        if (slot == NO_SLOT) continue;   // reported by cloud_eval_kernel
        double vin;
        if (p.in.records) vin = *reinterpret_cast<const double *>(p.in.words + (uint64_t) slot * p.in.stride + 8192);
        else vin = p.in.variance ? p.in.variance[slot] : p.default_var;
        var += p.var_w[e] * vin;
    }
    const uint32_t oslot = p.slot_of_row ? p.slot_of_row[r] : (uint32_t) r;
    if (p.out.records) {
        uint8_t *rec = p.out.words + (uint64_t) oslot * p.out.stride;
        reinterpret_cast<uint32_t *>(rec - 8)[0] = p.out_bidx[r];
        reinterpret_cast<int32_t *>(rec - 8)[1] = IDASH_B200_TLWE_SAMPLE_UID;
        *reinterpret_cast<double *>(rec + 8192) = var;
    } else {
        if (p.out.index) p.out.index[oslot] = p.out_bidx[r];
        if (p.out.variance) p.out.variance[oslot] = var;
    }
}

__global__ void slot_map_kernel(CtView in, uint32_t *slot_of_ct, uint32_t n_ct_slots) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.count) return;
    const uint32_t idx = in.records ? *reinterpret_cast<const uint32_t *>(in.words + i * in.stride - 8) : in.index[i];
    if (idx < n_ct_slots) slot_of_ct[idx] = (uint32_t) i;
}

// ------------------------------------------------------------------------------------------------
// K4: exact phase + decode
// ------------------------------------------------------------------------------------------------
struct KeyBits { uint32_t w[32]; };   // bit i of the binary TRLWE key

// One CTA (128 threads) per ciphertext; thread t owns phase coefficients j = 8t .. 8t+7.
// ext[d + 1024] = a[d] for d >= 0, -a[d + 1024] for d < 0 (the negacyclic extension), so that
//     phase[j] = b[j] - sum_{i : key_i = 1} ext[1024 + j - i].
// The i loop runs in steps of 8 over a 16-word register window that slides down by one aligned
// 8-word block per step; key bits are warp-uniform, so skipped bits cost one predicate test.
__global__ void __launch_bounds__(128) decrypt_kernel(CtView in, uint64_t n_ct, KeyBits key, uint32_t S, float *scores, uint32_t *phase) {
    __shared__ __align__(16) uint32_t ext[2 * POLY_N];
    const uint32_t t = threadIdx.x;
    for (uint64_t c = blockIdx.x; c < n_ct; c += gridDim.x) {
        const uint8_t *a = in.words + c * in.stride;
        const uint8_t *b = a + 4 * POLY_N;
        for (uint32_t i = 4 * t; i < POLY_N; i += 4 * 128) {
            const uint4 v = ldg128(a + 4 * i);
            *reinterpret_cast<uint4 *>(ext + POLY_N + i) = v;
            *reinterpret_cast<uint4 *>(ext + i) = make_uint4(0u - v.x, 0u - v.y, 0u - v.z, 0u - v.w);
        }
        __syncthreads();
        const uint32_t j0 = 8 * t;
        uint32_t acc[8];
        {
            const uint4 b0 = ldg128(b + 4 * j0), b1 = ldg128(b + 4 * j0 + 16);
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
            acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
        }
        uint32_t W[16];   // W[8..15] = ext[base .. base+7], W[0..7] = ext[base-8 .. base-1], base = 1024 + j0 - 8q
        {
            const uint4 h0 = *reinterpret_cast<const uint4 *>(ext + POLY_N + j0);
            const uint4 h1 = *reinterpret_cast<const uint4 *>(ext + POLY_N + j0 + 4);
            W[8] = h0.x; W[9] = h0.y; W[10] = h0.z; W[11] = h0.w; W[12] = h1.x; W[13] = h1.y; W[14] = h1.z; W[15] = h1.w;
        }
        for (uint32_t q = 0; q < POLY_N / 8; ++q) {
            const uint32_t base = POLY_N + j0 - 8 * q;
            const uint4 l0 = *reinterpret_cast<const uint4 *>(ext + base - 8);
            const uint4 l1 = *reinterpret_cast<const uint4 *>(ext + base - 4);
            W[0] = l0.x; W[1] = l0.y; W[2] = l0.z; W[3] = l0.w; W[4] = l1.x; W[5] = l1.y; W[6] = l1.z; W[7] = l1.w;
            const uint32_t bits = (key.w[q >> 2] >> ((q & 3u) * 8u)) & 0xFFu;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (bits & (1u << r)) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) acc[m] -= W[8 + m - r];
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) W[8 + m] = W[m];
        }
        if (phase) {
            uint4 *pp = reinterpret_cast<uint4 *>(phase + c * POLY_N + j0);
            pp[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            pp[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
        }
        if (scores) {
            // (float) (double(int32) / 2^32): one rounding of a 32-bit integer to 24 bits, then an exact
            // power-of-two scale -- identical to idash.cpp:718 + numeric-functions.cpp:36-38
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (j0 + m < S) scores[c * S + j0 + m] = __int2float_rn((int32_t) acc[m]) * 2.3283064365386963e-10f;
        }
        __syncthreads();
    }
}

#include "decrypt_tc.cuh"
#include "decrypt_pair.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return IDASH_B200_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CUDA_TRY(cudaMalloc(&p, bytes));
        cap = bytes;
        return IDASH_B200_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct idash_b200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;     // used by the *_host entry points
    int *d_status = nullptr;
    uint64_t *d_var_flag = nullptr;    // variance_scan_kernel's result (2 words)
    int *h_status = nullptr;           // pinned
    uint64_t launches = 0;
    int kernel_choice = IDASH_B200_KERNEL_AUTO;
    int last_kernel = 0;               // which cloud kernel the last launch used
    int decrypt_choice = IDASH_B200_DECRYPT_AUTO;
    int last_decrypt_kernel = 0;
    std::vector<cudaEvent_t> t_begin, t_end;   // per-launch timing of the dominant kernels (timing_enable)
    int t_used = 0;
    DevBuf in_buf, out_buf, slot_buf, row_slot_buf, aux_in_idx, aux_in_var, aux_out_idx, aux_out_var, scores_buf, phase_buf;
    // pipelined host path (cloud_eval_host): copy-in / copy-out streams and per-piece events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    // the small per-row finalize kernel runs beside the main kernel on its own stream (fork / join events)
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

// A launch over part of the model: tiles [tile_lo, tile_hi) of the ring kernel = rows [row_lo, row_hi) (caller rows are
// the sorted rows). `first`: also build the ciphertext-index -> slot table (once per evaluation).
struct Piece { uint32_t tile_lo, tile_hi; uint64_t row_lo, row_hi; bool first; };

struct idash_b200_model {
    idash_b200_layout *layout = nullptr;
    int device = 0;
    idash_b200_group *d_groups = nullptr;      // IMAD groups of the overflow rows (rows no tile holds)
    idash_b200_entry *d_entries = nullptr;
    mutable idash_b200_group *d_groups_full = nullptr;   // IMAD groups of every row: uploaded when the IMAD kernel evaluates the whole model
    mutable idash_b200_entry *d_entries_full = nullptr;
    uint64_t *d_var_ptr = nullptr;
    uint32_t *d_var_ct = nullptr;
    double *d_var_w = nullptr;
    double *d_var_wsum = nullptr;
    uint32_t *d_out_bidx = nullptr;
    idash_b200_tile *d_tiles = nullptr;      // tensor-core layout (null when not eligible)
    uint32_t *d_tile_rows = nullptr;
    int32_t *d_tile_bias = nullptr;
    uint8_t *d_tile_coef = nullptr;
    uint32_t *d_tile_used = nullptr;
    uint32_t *d_feat_used = nullptr;         // ring variant only
    mutable int rows_identity = -1;          // cached: tile t holds exactly the caller rows 64 t .. (pipelined host path)
};

extern "C" int idash_b200_init(idash_b200_ctx **out, int device) {
    clear_error();
    if (!out) return set_error(IDASH_B200_ERR_INVALID, "init: null argument");
    *out = nullptr;
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return set_error(IDASH_B200_ERR_CUDA, "init: device %d not available (%d CUDA devices)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    idash_b200_ctx *c = new (std::nothrow) idash_b200_ctx();
    if (!c) return set_error(IDASH_B200_ERR_NOMEM, "init: out of memory");
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&c->d_status, sizeof(int)));
    CUDA_TRY(cudaMalloc(&c->d_var_flag, 2 * sizeof(uint64_t)));
    CUDA_TRY(cudaMemset(c->d_status, 0, sizeof(int)));
    CUDA_TRY(cudaMallocHost(&c->h_status, sizeof(int)));
    const int tc_smem_max = (int) tc_smem_bytes(IDASH_B200_TILE_KMAX);
    CUDA_TRY(cudaFuncSetAttribute(cloud_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_max));
    CUDA_TRY(cudaFuncSetAttribute(cloud_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_max));
    CUDA_TRY(cudaFuncSetAttribute(cloud_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_max));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_tc_kernel<IDASH_B200_CT_BYTES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_tc_kernel<IDASH_B200_CT_BYTES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_tc_kernel<IDASH_B200_RECORD_BYTES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_tc_kernel<IDASH_B200_RECORD_BYTES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_pair_kernel<IDASH_B200_CT_BYTES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_pair_kernel<IDASH_B200_CT_BYTES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_pair_kernel<IDASH_B200_RECORD_BYTES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(decrypt_pair_kernel<IDASH_B200_RECORD_BYTES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    for (int rot = 0; rot < 2; ++rot)
        for (int bat = 0; bat < 2; ++bat)
            for (int epi = 0; epi < RG_EPI_VARIANTS; ++epi)
                CUDA_TRY(cudaFuncSetAttribute(ring_kernel_fn(rot != 0, bat != 0, epi), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) RG_SMEM_MAX));
    *out = c;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_destroy(idash_b200_ctx *c) {
    if (!c) return IDASH_B200_OK;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_k) cudaEventDestroy(e);
    DevBuf *bufs[] = {&c->in_buf, &c->out_buf, &c->slot_buf, &c->row_slot_buf, &c->aux_in_idx, &c->aux_in_var,
                      &c->aux_out_idx, &c->aux_out_var, &c->scores_buf, &c->phase_buf};
    for (DevBuf *b : bufs) b->release();
    for (cudaEvent_t e : c->t_begin) cudaEventDestroy(e);
    for (cudaEvent_t e : c->t_end) cudaEventDestroy(e);
    if (c->d_status) cudaFree(c->d_status);
    if (c->d_var_flag) cudaFree(c->d_var_flag);
    if (c->h_status) cudaFreeHost(c->h_status);
    delete c;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_timing_enable(idash_b200_ctx *c, int max_launches) {
    clear_error();
    if (!c || max_launches < 0) return set_error(IDASH_B200_ERR_INVALID, "timing_enable: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    for (cudaEvent_t e : c->t_begin) cudaEventDestroy(e);
    for (cudaEvent_t e : c->t_end) cudaEventDestroy(e);
    c->t_begin.clear(); c->t_end.clear(); c->t_used = 0;
    for (int i = 0; i < max_launches; ++i) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a));
        CUDA_TRY(cudaEventCreate(&b));
        c->t_begin.push_back(a); c->t_end.push_back(b);
    }
    return IDASH_B200_OK;
}

extern "C" int idash_b200_timing_read(idash_b200_ctx *c, float *ms, int *n) {
    clear_error();
    if (!c || !ms || !n) return set_error(IDASH_B200_ERR_INVALID, "timing_read: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const int cnt = std::min(*n, c->t_used);
    for (int i = 0; i < cnt; ++i) {
        CUDA_TRY(cudaEventSynchronize(c->t_end[i]));
        CUDA_TRY(cudaEventElapsedTime(&ms[i], c->t_begin[i], c->t_end[i]));
    }
    *n = cnt;
    c->t_used = 0;
    return IDASH_B200_OK;
}

extern "C" uint64_t idash_b200_kernel_launches(const idash_b200_ctx *c) { return c ? c->launches : 0; }

extern "C" int idash_b200_set_kernel(idash_b200_ctx *c, int which) {
    clear_error();
    if (!c || which < IDASH_B200_KERNEL_AUTO || which > IDASH_B200_KERNEL_TENSOR_RING) return set_error(IDASH_B200_ERR_INVALID, "set_kernel: bad argument");
    c->kernel_choice = which;
    return IDASH_B200_OK;
}
extern "C" int idash_b200_last_kernel(const idash_b200_ctx *c) { return c ? c->last_kernel : 0; }

extern "C" int idash_b200_set_decrypt_kernel(idash_b200_ctx *c, int which) {
    clear_error();
    if (!c || which < IDASH_B200_DECRYPT_AUTO || which > IDASH_B200_DECRYPT_TENSOR_PAIR) return set_error(IDASH_B200_ERR_INVALID, "set_decrypt_kernel: bad argument");
    c->decrypt_choice = which;
    return IDASH_B200_OK;
}
extern "C" int idash_b200_last_decrypt_kernel(const idash_b200_ctx *c) { return c ? c->last_decrypt_kernel : 0; }

extern "C" int idash_b200_host_alloc(void **ptr, size_t bytes) {
    clear_error();
    if (!ptr) return set_error(IDASH_B200_ERR_INVALID, "host_alloc: null argument");
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
    return IDASH_B200_OK;
}
extern "C" int idash_b200_host_register(void *ptr, size_t bytes) {
    clear_error();
    if (!ptr || !bytes) return set_error(IDASH_B200_ERR_INVALID, "host_register: null argument");
    CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return IDASH_B200_OK;
}
extern "C" int idash_b200_host_unregister(void *ptr) {
    clear_error();
    if (ptr) CUDA_TRY(cudaHostUnregister(ptr));
    return IDASH_B200_OK;
}
extern "C" int idash_b200_host_free(void *ptr) {
    clear_error();
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return IDASH_B200_OK;
}

template <typename T>
static int upload(T **dst, const T *src, size_t n) {
    *dst = nullptr;
    CUDA_TRY(cudaMalloc((void **) dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return IDASH_B200_OK;
}

extern "C" int idash_b200_model_free(idash_b200_model *m) {
    if (!m) return IDASH_B200_OK;
    cudaSetDevice(m->device);
    cudaFree(m->d_groups); cudaFree(m->d_entries); cudaFree(m->d_groups_full); cudaFree(m->d_entries_full); cudaFree(m->d_var_ptr); cudaFree(m->d_var_ct); cudaFree(m->d_var_w); cudaFree(m->d_var_wsum);
    cudaFree(m->d_out_bidx);
    cudaFree(m->d_tiles); cudaFree(m->d_tile_rows); cudaFree(m->d_tile_bias); cudaFree(m->d_tile_coef); cudaFree(m->d_tile_used);
    cudaFree(m->d_feat_used);
    idash_b200_layout_free(m->layout);
    delete m;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_model_upload_layout(idash_b200_ctx *c, idash_b200_layout *L, idash_b200_model **out) {
    clear_error();
    if (!c || !L || !out) { idash_b200_layout_free(L); return set_error(IDASH_B200_ERR_INVALID, "model_upload_layout: null argument"); }
    *out = nullptr;
    idash_b200_model *m = new (std::nothrow) idash_b200_model();
    if (!m) { idash_b200_layout_free(L); return set_error(IDASH_B200_ERR_NOMEM, "model_upload: out of memory"); }
    m->layout = L;
    m->device = c->device;
    int rc = IDASH_B200_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = set_error(IDASH_B200_ERR_CUDA, "model_upload: cudaSetDevice(%d) failed", c->device);
    if (rc ||
        (rc = upload(&m->d_var_wsum, L->var_wsum.data(), L->var_wsum.size())) ||
        (rc = upload(&m->d_var_ptr, L->var_ptr.data(), L->var_ptr.size())) ||
        (rc = upload(&m->d_var_ct, L->var_ct.data(), L->var_ct.size())) ||
        (rc = upload(&m->d_var_w, L->var_w.data(), L->var_w.size())) ||
        (rc = upload(&m->d_out_bidx, L->out_bidx.data(), L->out_bidx.size())) ||
        (!L->groups.empty() && ((rc = upload(&m->d_groups, L->groups.data(), L->groups.size())) ||
                                (rc = upload(&m->d_entries, L->entries.data(), L->entries.size())))) ||
        (L->groups_all && ((rc = upload(&m->d_groups_full, L->groups_full.data(), L->groups_full.size())) ||
                           (rc = upload(&m->d_entries_full, L->entries_full.data(), L->entries_full.size())))) ||
        (!L->tiles.empty() &&
         ((rc = upload(&m->d_tiles, L->tiles.data(), L->tiles.size())) ||
          (rc = upload(&m->d_tile_rows, L->tile_rows.data(), L->tile_rows.size())) ||
          (rc = upload(&m->d_tile_bias, L->tile_bias.data(), L->tile_bias.size())) ||
          (rc = upload(&m->d_tile_coef, L->tile_coef.data(), L->tile_coef.size())) ||
          (rc = upload(&m->d_tile_used, L->tile_used.data(), L->tile_used.size())) ||
          (L->ring_ok && (rc = upload(&m->d_feat_used, L->feat_used.data(), L->feat_used.size())))))) {
        idash_b200_model_free(m);
        return rc;
    }
    *out = m;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_model_clone(idash_b200_ctx *c, const idash_b200_model *src, idash_b200_model **out) {
    clear_error();
    if (!c || !src || !out) return set_error(IDASH_B200_ERR_INVALID, "model_clone: null argument");
    src->layout->refs.fetch_add(1);
    return idash_b200_model_upload_layout(c, src->layout, out);
}

extern "C" int idash_b200_model_upload(idash_b200_ctx *c, const idash_b200_model_desc *desc, idash_b200_model **out) {
    clear_error();
    if (!c || !desc || !out) return set_error(IDASH_B200_ERR_INVALID, "model_upload: null argument");
    *out = nullptr;
    idash_b200_layout *L = nullptr;
    // the IMAD groups of every row are only built when that kernel has been chosen for the whole model (or later, on demand)
    const uint32_t flags = c->kernel_choice == IDASH_B200_KERNEL_IMAD ? IDASH_B200_COMPILE_GROUPS_ALL : IDASH_B200_COMPILE_DEFAULT;
    int rc = idash_b200_layout_compile_ex(desc, flags, &L);
    if (rc) return rc;
    return idash_b200_model_upload_layout(c, L, out);
}

// IMAD groups of every row on the device (built and uploaded on first use)
static int ensure_full_groups(const idash_b200_model *m) {
    if (m->d_groups_full) return IDASH_B200_OK;
    int rc = idash_b200_layout_ensure_groups_all(m->layout);
    if (rc) return rc;
    if ((rc = upload(&m->d_groups_full, m->layout->groups_full.data(), m->layout->groups_full.size()))) return rc;
    return upload(&m->d_entries_full, m->layout->entries_full.data(), m->layout->entries_full.size());
}

// The IMAD kernel over a list of groups
static int launch_imad(idash_b200_ctx *c, const idash_b200_layout *L, const idash_b200_group *d_groups, const idash_b200_entry *d_entries,
                       uint64_t n_groups, const CtView &in, const CtView &out, const uint32_t *d_slot_of_ct, uint32_t n_ct_slots,
                       const uint32_t *d_slot_of_row, cudaStream_t st) {
    if (n_groups == 0) return IDASH_B200_OK;
    const int mode = (L->NR == 1) ? 0 : (L->shifts_aligned ? 1 : 2);
    CloudParams p;
    memset(&p, 0, sizeof(p));
    p.groups = d_groups;
    p.entries = d_entries;
    p.n_groups = (uint32_t) n_groups;
    p.in = in;
    p.out = out;
    p.slot_of_ct = d_slot_of_ct;
    p.n_ct_slots = n_ct_slots;
    p.slot_of_row = d_slot_of_row;
    p.S = L->S;
    p.RS = L->RS;
    p.status = c->d_status;
    // enough CTAs for >= ~8 waves of 8 resident CTAs/SM, at most 16 consecutive groups per CTA
    uint32_t gpc = 16;
    while (gpc > 1 && (uint64_t) ((p.n_groups + gpc - 1) / gpc) * 4 < (uint64_t) c->sm_count * 64) gpc >>= 1;
    p.groups_per_cta = gpc;
    const unsigned grid = ((p.n_groups + gpc - 1) / gpc) * 4;
    if (mode == 0) cloud_eval_kernel<0><<<grid, 128, 0, st>>>(p);
    else if (mode == 1) cloud_eval_kernel<1><<<grid, 128, 0, st>>>(p);
    else cloud_eval_kernel<2><<<grid, 128, 0, st>>>(p);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return IDASH_B200_OK;
}

extern "C" int idash_b200_model_get_info(const idash_b200_model *m, idash_b200_model_info *info) {
    clear_error();
    if (!m) return set_error(IDASH_B200_ERR_INVALID, "model_get_info: null argument");
    return idash_b200_layout_get_info(m->layout, info);
}

// Builds the device view of a ciphertext array that already lives in device memory.
static int make_view(const idash_b200_cts *a, bool is_output, CtView *v, const char *what) {
    if (!a) return set_error(IDASH_B200_ERR_INVALID, "%s: null ciphertext array", what);
    if (a->count && !a->data) return set_error(IDASH_B200_ERR_INVALID, "%s: null data pointer", what);
    memset(v, 0, sizeof(*v));
    v->count = a->count;
    if (a->layout == IDASH_B200_LAYOUT_PACKED) {
        if ((uintptr_t) a->data & 15u) return set_error(IDASH_B200_ERR_INVALID, "%s: packed data must be 16-byte aligned", what);
        v->words = (uint8_t *) a->data;
        v->stride = IDASH_B200_CT_BYTES;
        v->index = a->index;
        v->variance = a->variance;
        v->records = 0;
    } else if (a->layout == IDASH_B200_LAYOUT_RECORDS) {
        if (a->count && ((uintptr_t) a->data & 15u) != 8u)
            return set_error(IDASH_B200_ERR_INVALID, "%s: record stream must start at an address = 8 mod 16 (file image + 8)", what);
        if (a->index || a->variance) return set_error(IDASH_B200_ERR_INVALID, "%s: index/variance must be NULL for the RECORDS layout", what);
        v->words = (uint8_t *) a->data + 8;
        v->stride = IDASH_B200_RECORD_BYTES;
        v->records = 1;
    } else {
        return set_error(IDASH_B200_ERR_INVALID, "%s: unknown layout %d", what, a->layout);
    }
    (void) is_output;
    return IDASH_B200_OK;
}

// Launch geometry of the persistent ring kernel for a model on this device (shared by the eligibility checks and the launch, so
// that a shape whose shared-memory budget does not work out selects another kernel / per-batch launches instead of failing).
struct RingPlan { uint32_t n_slices, n_chunks, n_slots, n_bchunks, max_chunk_tiles, extra_tiles; };
static bool ring_plan(const idash_b200_ctx *c, const idash_b200_layout *L, uint64_t n_tiles, uint32_t n_batches, RingPlan *rp);
static bool ring_selected(const idash_b200_ctx *c, const idash_b200_layout *L);

// Debug / tuning environment variables are only honoured by the profiling build (-DIDASH_B200_PROFILE, libidash_b200_prof.so):
// a stray variable must not change what the production library computes or launches.
static inline const char *debug_env(const char *name) {
#ifdef IDASH_B200_PROFILE
    return getenv(name);
#else
    (void) name;
    return nullptr;
#endif
}

// n_batches > 1 (batched launch): ins / outs hold the views of every batch (in == ins[0], out == outs[0]); the caller has checked
// that the ring kernel takes the model and that the inputs are in identity order.
static int launch_cloud(idash_b200_ctx *c, const idash_b200_model *m, const CtView &in, const CtView &out,
                        const uint32_t *d_slot_of_row, cudaStream_t st, const Piece *piece = nullptr,
                        uint32_t n_batches = 1, const CtView *ins = nullptr, const CtView *outs = nullptr,
                        const idash_b200_model *const *bms = nullptr) {     // bms: the model of every batch (null: all batches share m)
    const idash_b200_layout *L = m->layout;
    if (out.count != L->n_rows) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: out->count (%llu) != model rows (%llu)",
                                                 (unsigned long long) out.count, (unsigned long long) L->n_rows);
    if (L->n_rows == 0) return IDASH_B200_OK;
    if (in.count >= NO_SLOT) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: too many input ciphertexts");

    // ciphertext index -> input slot
    const bool has_entries = L->ct_min <= L->ct_max;
    const uint32_t *d_slot_of_ct = nullptr;
    uint32_t n_ct_slots = (uint32_t) in.count;
    const bool identity = !in.records && in.index == nullptr;
    if (!identity) {
        n_ct_slots = has_entries ? L->ct_max + 1 : 1;
        int rc = c->slot_buf.ensure((size_t) n_ct_slots * 4);
        if (rc) return rc;
        if (!piece || piece->first) {
            CUDA_TRY(cudaMemsetAsync(c->slot_buf.p, 0xFF, (size_t) n_ct_slots * 4, st));
            if (in.count) {
                slot_map_kernel<<<(unsigned) ((in.count + 255) / 256), 256, 0, st>>>(in, (uint32_t *) c->slot_buf.p, n_ct_slots);
                c->launches++;
            }
        }
        d_slot_of_ct = (const uint32_t *) c->slot_buf.p;
    }

    const int mode = (L->NR == 1) ? 0 : (L->shifts_aligned ? 1 : 2);
    const bool tc_ok = !L->tiles.empty();
    if (c->kernel_choice >= IDASH_B200_KERNEL_TENSOR && !tc_ok)
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: the tensor-core kernel was requested but no row of the model is eligible "
                                                 "(coefficients outside int16 or windows wider than %u features)", IDASH_B200_RING_KMAX);
    if (c->kernel_choice == IDASH_B200_KERNEL_TENSOR_RING && !L->ring_ok)
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: the persistent ring kernel was requested but the model is not eligible "
                                                 "(needs forward-moving bands of at most %u features)", IDASH_B200_RING_KMAX);
    const bool use_tc = tc_ok && c->kernel_choice != IDASH_B200_KERNEL_IMAD;
    if (c->kernel_choice == IDASH_B200_KERNEL_TENSOR_RING && L->ring_ok && !ring_selected(c, L))
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: the persistent ring kernel was requested but the model has too many tiles per chunk");
    const bool use_ring = ring_selected(c, L);
    if (piece && (!use_ring || L->n_overflow_rows)) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: internal: partial launches need the ring kernel");
    if (!use_tc && L->n_overflow_rows != L->n_rows) { int rc = ensure_full_groups(m); if (rc) return rc; }
    // rows no tile holds (an outlier window, a coefficient outside the limb range): the IMAD kernel evaluates them beside the
    // tensor-core kernel -- queued first, it is a few CTAs that finish while the persistent grid ramps up
    if (use_tc && L->n_overflow_rows) {
        int rc = launch_imad(c, L, m->d_groups, m->d_entries, L->groups.size(), in, out, d_slot_of_ct, n_ct_slots, d_slot_of_row, st);
        if (rc) return rc;
    }
    // Per-row variance / record headers: independent of the main kernel's words, so it is forked onto its own stream
    // (after the slot table is ready on `st`) and joined at the end -- it hides behind the main kernel.
    if (!c->s_aux) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(c->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
    FinalizeParams f;
    memset(&f, 0, sizeof(f));
    f.row_lo = piece ? piece->row_lo : 0;
    f.n_rows = piece ? piece->row_hi : L->n_rows;
    f.var_ptr = m->d_var_ptr;
    f.var_ct = m->d_var_ct;
    f.var_w = m->d_var_w;
    f.out_bidx = m->d_out_bidx;
    f.in = in;
    f.out = out;
    f.slot_of_ct = d_slot_of_ct;
    f.n_ct_slots = n_ct_slots;
    f.slot_of_row = d_slot_of_row;
    f.default_var = 8.8817841970012523e-16;   // alpha^2 = 2^-50 (eval/idash.cpp:20, tlwe-functions.cpp:38)
    f.var_wsum = m->d_var_wsum;
    for (uint32_t b = 0; b < n_batches; ++b) {
        if (n_batches > 1) { f.in = ins[b]; f.out = outs[b]; }
        if (bms) { f.var_ptr = bms[b]->d_var_ptr; f.var_ct = bms[b]->d_var_ct; f.var_w = bms[b]->d_var_w; f.out_bidx = bms[b]->d_out_bidx; f.var_wsum = bms[b]->d_var_wsum; }
        f.var_uniform = nullptr;
        if ((f.in.records || f.in.variance) && f.in.count) {
            CUDA_TRY(cudaMemsetAsync(c->d_var_flag, 0, 2 * sizeof(uint64_t), c->s_aux));
            variance_scan_kernel<<<(unsigned) ((f.in.count + 255) / 256), 256, 0, c->s_aux>>>(f.in, c->d_var_flag);
            c->launches++;
            f.var_uniform = c->d_var_flag;
        }
        cloud_finalize_kernel<<<(unsigned) ((f.n_rows - f.row_lo + 63) / 64), 64, 0, c->s_aux>>>(f);   // 64-thread CTAs fit beside a resident ring CTA
        c->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev_join, c->s_aux));

    const bool timed = c->t_used < (int) c->t_begin.size();
    if (timed) CUDA_TRY(cudaEventRecord(c->t_begin[c->t_used], st));
    if (use_ring) {
        RingParams p;
        memset(&p, 0, sizeof(p));
        p.tiles = m->d_tiles; p.tile_rows = m->d_tile_rows; p.tile_bias = m->d_tile_bias; p.tile_coef = m->d_tile_coef;
        p.feat_used = m->d_feat_used;
        p.n_feat_words = (uint32_t) L->feat_used.size();
        p.n_tiles = piece ? piece->tile_hi - piece->tile_lo : (uint32_t) L->tiles.size();
        p.tile_base = piece ? piece->tile_lo : 0u;
        RingPlan rp;
        if (!ring_plan(c, L, p.n_tiles, n_batches, &rp)) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: internal: ring kernel shared-memory budget");
        p.n_slices = rp.n_slices; p.n_chunks = rp.n_chunks; p.n_slots = rp.n_slots; p.n_bchunks = rp.n_bchunks; p.max_chunk_tiles = rp.max_chunk_tiles;
        p.extra_tiles = rp.extra_tiles;
        const uint32_t max_nb = L->tile_kmax / 32u;
        p.n_batches = n_batches;
        for (uint32_t b = 0; b < n_batches && n_batches > 1; ++b) {
            const idash_b200_model *mb = bms ? bms[b] : m;
            p.batch_in[b] = (unsigned long long) ins[b].words; p.batch_out[b] = (unsigned long long) outs[b].words;
            p.batch_tiles[b] = mb->d_tiles; p.batch_rows[b] = mb->d_tile_rows; p.batch_bias[b] = mb->d_tile_bias; p.batch_coef[b] = mb->d_tile_coef;
            p.batch_feat_used[b] = mb->d_feat_used; p.batch_nfw[b] = (uint32_t) mb->layout->feat_used.size(); p.batch_S[b] = mb->layout->S;
            p.batch_coef_bytes[b] = mb->layout->tile_coef.size();
        }
        p.meta_off = p.n_slots * RG_BLOCK_BYTES + p.n_bchunks * TC_B_CHUNK;
        p.zero_off = p.meta_off + RG_META_BYTES;
        p.hdr_off = p.zero_off + (L->NR != 1 ? RG_ZERO_BYTES : 0u);
        p.in = in; p.out = out;
        p.slot_of_ct = d_slot_of_ct; p.n_ct_slots = n_ct_slots; p.slot_of_row = d_slot_of_row;
        p.S = L->S; p.NR = L->NR; p.RS = L->RS;
        p.coef_bytes = L->tile_coef.size();
        p.coef_prefetch = max_nb >= 5u ? 2u : 0u;     // measured: 0.673 -> 0.640 ms at neighbors = 50 (7 blocks), nothing to gain on narrow bands
        if (const char *cp = debug_env("IDASH_B200_COEF_PREFETCH")) p.coef_prefetch = (uint32_t) atoi(cp);
        p.status = c->d_status;
        if (const char *ko = debug_env("IDASH_B200_KNOCKOUT")) p.knockout = (uint32_t) atoi(ko);
        p.tune = L->NR != 1 ? RG_TUNE_ROT : RG_TUNE_DEFAULT;
        if (const char *tu = debug_env("IDASH_B200_TUNE")) p.tune = (uint32_t) atoi(tu);
        if (const char *tr = debug_env("IDASH_B200_TRACE")) p.trace_cta = (uint32_t) atoi(tr) + 1u;
        const size_t ring_smem = ring_smem_bytes(p.n_slots, p.n_bchunks, p.max_chunk_tiles, L->NR != 1);
        const dim3 grid(p.n_slices * (p.n_chunks + (p.extra_tiles ? 1u : 0u)));
        int epi = RG_EPI_DEFAULT;
#ifdef IDASH_B200_PROFILE
        if (debug_env("IDASH_B200_TUNE")) epi = (p.tune & 2048u) ? 3 : (p.tune & 512u) ? 2 : (p.tune & 256u) ? 1 : 0;
#endif
        ring_kernel_fn(L->NR != 1, n_batches > 1, epi)<<<grid, RG_THREADS, ring_smem, st>>>(p);
        if (p.trace_cta) {   // debugging only: dump the timeline of the traced CTA to the file named by IDASH_B200_TRACE_FILE
            static unsigned long long h[RG_TRACE_TILES * RG_TRACE_EVENTS];
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaMemcpyFromSymbol(h, g_ring_trace, sizeof(h)));
            const char *fn = debug_env("IDASH_B200_TRACE_FILE");
            if (FILE *f = fopen(fn ? fn : "ring_trace.txt", "w")) {
                for (int i = 0; i < RG_TRACE_TILES; ++i) {
                    for (int e = 0; e < RG_TRACE_EVENTS; ++e) fprintf(f, "%llu ", h[i * RG_TRACE_EVENTS + e]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
        c->last_kernel = IDASH_B200_KERNEL_TENSOR_RING;
    } else if (use_tc) {
        TcParams p;
        memset(&p, 0, sizeof(p));
        p.tiles = m->d_tiles; p.tile_rows = m->d_tile_rows; p.tile_bias = m->d_tile_bias; p.tile_coef = m->d_tile_coef;
        p.tile_used = m->d_tile_used;
        p.n_tiles = (uint32_t) L->tiles.size();
        p.in = in; p.out = out;
        p.slot_of_ct = d_slot_of_ct; p.n_ct_slots = n_ct_slots; p.slot_of_row = d_slot_of_row;
        p.S = L->S; p.NR = L->NR; p.RS = L->RS;
        p.status = c->d_status;
        if ((uint64_t) p.n_tiles * 16 > 0x7FFFFFFFull) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval: too many tiles");
        const unsigned grid = p.n_tiles * 16u;
        const size_t smem = tc_smem_bytes(L->tile_kmax);
        if (mode == 0) cloud_tc_kernel<0><<<grid, TC_THREADS, smem, st>>>(p);
        else if (mode == 1) cloud_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(p);
        else cloud_tc_kernel<2><<<grid, TC_THREADS, smem, st>>>(p);
        c->last_kernel = IDASH_B200_KERNEL_TENSOR_TILE;
    } else {
        const bool all_overflow = L->n_overflow_rows == L->n_rows;
        int rc = launch_imad(c, L, all_overflow ? m->d_groups : m->d_groups_full, all_overflow ? m->d_entries : m->d_entries_full,
                             all_overflow ? L->groups.size() : L->groups_full.size(), in, out, d_slot_of_ct, n_ct_slots, d_slot_of_row, st);
        if (rc) return rc;
        c->launches--;     // counted below
        c->last_kernel = IDASH_B200_KERNEL_IMAD;
    }
    c->launches++;
    if (timed) CUDA_TRY(cudaEventRecord(c->t_end[c->t_used++], st));
    CUDA_TRY(cudaGetLastError());

    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));     // join: the launch is complete when both kernels are
    CUDA_TRY(cudaGetLastError());
    return IDASH_B200_OK;
}

static bool ring_plan(const idash_b200_ctx *c, const idash_b200_layout *L, uint64_t n_tiles, uint32_t n_batches, RingPlan *rp) {
    // NUM_REGIONS > 1: b[RS..N) of every output is zero (eval/idash.cpp:839-841), so only the 8 slices of polynomial a and the
    // ceil(RS / 128) slices of b that hold kept words are computed; the CTAs zero-fill the rest. Fewer slices = more chunks = fewer
    // tiles per CTA (NUM_REGIONS = 3: 11 x 13 CTAs instead of 16 x 9).
    rp->n_slices = L->NR == 1 ? 16u : 8u + (L->RS + 127u) / 128u;
    if (const char *ns = debug_env("IDASH_B200_RING_SLICES")) rp->n_slices = std::max(rp->n_slices, std::min(16u, (uint32_t) atoi(ns)));   // experiments
    if ((uint32_t) c->sm_count < rp->n_slices || n_tiles == 0 || L->tile_kmax == 0) return false;
    rp->n_chunks = (uint32_t) c->sm_count / rp->n_slices;
    // SMs left over by that division (148 = 16 x 9 + 4, or 11 x 13 + 5) take one more, short chunk of X tiles at the end of the list: its
    // n_slices CTAs run on the spare SMs in `waves` waves, each paying the pipeline's ramp-up (counted as `fill` tiles), so
    // waves (X + fill) = R, the regular chunk's length, and n_chunks R + X = tiles
    uint32_t n_spare = (uint32_t) c->sm_count - rp->n_slices * rp->n_chunks, fill = 4;     // fill: measured 2 .. 12, profiles/r02_ring_extra_chunk.txt
    rp->extra_tiles = 0;
    if (const char *nc = debug_env("IDASH_B200_RING_CHUNKS")) { rp->n_chunks = std::max(1u, std::min(rp->n_chunks, (uint32_t) atoi(nc))); n_spare = 0; }   // experiments
    if (const char *ne = debug_env("IDASH_B200_RING_EXTRA")) n_spare = std::min(n_spare, (uint32_t) atoi(ne));
    if (const char *ef = debug_env("IDASH_B200_RING_EXTRA_FILL")) fill = (uint32_t) atoi(ef);
    if (n_spare) {
        const uint64_t waves = (rp->n_slices + n_spare - 1u) / n_spare, total = n_tiles * n_batches;
        const uint64_t lost = (uint64_t) rp->n_chunks * waves * fill;
        rp->extra_tiles = total > lost ? (uint32_t) ((total - lost) / (rp->n_chunks * waves + 1u)) : 0u;
    }
    const uint64_t chunk_tiles = (n_tiles * n_batches + rp->n_chunks - 1u) / rp->n_chunks + 1u;
    if (chunk_tiles * 4u > 65536u) return false;
    rp->max_chunk_tiles = (uint32_t) chunk_tiles;
    const uint32_t max_nb = L->tile_kmax / 32u;
    const auto chunks_for = [&](uint32_t slots) -> uint32_t {
        const uint32_t used = slots * RG_BLOCK_BYTES + RG_META_BYTES + (L->NR != 1 ? RG_ZERO_BYTES : 0u) + 4u * rp->max_chunk_tiles;
        return used >= RG_SMEM_MAX ? 0u : std::min<uint32_t>(64u, (RG_SMEM_MAX - used) / TC_B_CHUNK);
    };
    // Input-block slots: the widest tile's blocks + 2 (one being staged, one ahead), and at least two and a half tiles of coefficient
    // chunks beside them (the loader admits an image chunk by chunk, so that is as good as three). Measured per workload (A/B runs,
    // profiles/r02_ring_slots.txt): MORE slots are slower -- producers that run far ahead of the MMAs only add read traffic in front
    // of the stores: neighbors = 5: 9 slots 0.4577 ms, 5 slots 0.4538; 20: 9 slots 0.4766, 6 slots 0.4692; 50: 9 slots 0.590, 8 slots 0.553.
    const uint32_t lo = max_nb + 1u, hi = std::min<uint32_t>(RG_MAX_SLOTS, max_nb + 2u);
    uint32_t slots = 0;
    for (uint32_t s = hi; s >= lo && !slots; --s) if (chunks_for(s) >= 2u * max_nb + max_nb / 2u) slots = s;
    for (uint32_t s = hi; s >= lo && !slots; --s) if (chunks_for(s) >= 2u * max_nb) slots = s;
    if (!slots) return false;
    if (const char *rs = debug_env("IDASH_B200_RING_SLOTS")) {   // experiments
        const uint32_t s = (uint32_t) atoi(rs);
        if (s >= lo && s <= RG_MAX_SLOTS && chunks_for(s) >= 2u * max_nb) slots = s;
    }
    rp->n_slots = slots;
    rp->n_bchunks = chunks_for(slots);
    if (const char *bc = debug_env("IDASH_B200_RING_BCHUNKS")) rp->n_bchunks = std::max(2u * max_nb, std::min(rp->n_bchunks, (uint32_t) atoi(bc)));
    return true;
}

// Would launch_cloud pick the persistent ring kernel for this model under the ctx's kernel choice?
static bool ring_selected(const idash_b200_ctx *c, const idash_b200_layout *L) {
    if (L->tiles.empty() || !L->ring_ok) return false;
    if (c->kernel_choice == IDASH_B200_KERNEL_IMAD || c->kernel_choice == IDASH_B200_KERNEL_TENSOR_TILE) return false;
    RingPlan rp;
    // the ring kernel keeps a 4-byte header per tile of a chunk in shared memory (first block in 20 bits)
    return ring_plan(c, L, L->tiles.size(), 1, &rp) && (L->tiles.back().f_base >> 5) + IDASH_B200_TILE_KMAX / 32u < (1u << RG_HDR_A_BITS);
}

extern "C" int idash_b200_cloud_eval_device(idash_b200_ctx *c, const idash_b200_model *m, const idash_b200_cts *in,
                                            const idash_b200_cts *out, const uint32_t *slot_of_row, void *stream) {
    clear_error();
    if (!c || !m) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device: null argument");
    if (m->device != c->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device: model lives on device %d, ctx on %d", m->device, c->device);
    CUDA_TRY(cudaSetDevice(c->device));
    CtView vin, vout;
    int rc;
    if ((rc = make_view(in, false, &vin, "cloud_eval_device(in)"))) return rc;
    if ((rc = make_view(out, true, &vout, "cloud_eval_device(out)"))) return rc;
    return launch_cloud(c, m, vin, vout, slot_of_row, (cudaStream_t) stream);
}

// The same model on several input / output sets in ONE launch of the persistent ring kernel (eval/idash.cpp:763-848 once per
// set). This is what a GPU that evaluates its target range for several sample batches wants: N launches over 1/N of the targets
// each pay the kernel's ramp-up and tail N times. Sets the ring kernel cannot take in one launch are evaluated one after another.
extern "C" int idash_b200_cloud_eval_device_batched(idash_b200_ctx *c, const idash_b200_model *m, uint32_t n_batches,
                                                    const idash_b200_cts *in, const idash_b200_cts *out, void *stream) {
    clear_error();
    if (!c || !m || !in || !out) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_batched: null argument");
    if (n_batches == 0) return IDASH_B200_OK;
    if (m->device != c->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_batched: model lives on device %d, ctx on %d", m->device, c->device);
    CUDA_TRY(cudaSetDevice(c->device));
    const idash_b200_layout *L = m->layout;
    std::vector<CtView> vin(n_batches), vout(n_batches);
    int rc;
    bool one_launch = n_batches > 1 && n_batches <= RG_MAX_BATCHES && ring_selected(c, L) && L->n_overflow_rows == 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        if ((rc = make_view(&in[b], false, &vin[b], "cloud_eval_device_batched(in)"))) return rc;
        if ((rc = make_view(&out[b], true, &vout[b], "cloud_eval_device_batched(out)"))) return rc;
        // one launch: every set in the same layout and size, inputs in identity order (slot = ciphertext index)
        one_launch = one_launch && !vin[b].records && vin[b].index == nullptr && vin[b].count == vin[0].count && vin[b].stride == vin[0].stride &&
                     (vin[b].variance != nullptr) == (vin[0].variance != nullptr) &&
                     vout[b].records == vout[0].records && vout[b].count == vout[0].count && vout[b].stride == vout[0].stride &&
                     (vout[b].index != nullptr) == (vout[0].index != nullptr) && (vout[b].variance != nullptr) == (vout[0].variance != nullptr);
    }
    if (one_launch) {
        RingPlan rp;
        const uint64_t blocks = L->tiles.empty() ? 0 : (L->tiles.back().f_base >> 5) + IDASH_B200_TILE_KMAX / 32u;
        one_launch = ring_plan(c, L, L->tiles.size(), n_batches, &rp) && blocks < (1u << RG_BATCH_SHIFT);
    }
    if (one_launch) return launch_cloud(c, m, vin[0], vout[0], nullptr, (cudaStream_t) stream, nullptr, n_batches, vin.data(), vout.data());
    for (uint32_t b = 0; b < n_batches; ++b)
        if ((rc = launch_cloud(c, m, vin[b], vout[b], nullptr, (cudaStream_t) stream))) return rc;
    return IDASH_B200_OK;
}

// Several MODELS of the same shape on their own input / output sets in ONE launch of the persistent ring kernel: BASELINE configs[3],
// the population-stratified model sets (_AFR / _AMR / _EUR: three `cloud` runs of the reference, eval/idash.cpp:763-848 once per
// population). The launch walks the tiles of every (model, input set) pair as virtual tiles, like the batched launch of one model;
// every pair brings its own tile list, coefficient images, row / Constant tables and NUM_SAMPLES. Falls back to one launch per pair
// when the models do not have the same shape or the ring kernel does not take them.
extern "C" int idash_b200_cloud_eval_device_multi_model(idash_b200_ctx *c, uint32_t n, const idash_b200_model *const *ms,
                                                        const idash_b200_cts *in, const idash_b200_cts *out, void *stream) {
    clear_error();
    if (!c || !ms || !in || !out) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_multi_model: null argument");
    if (n == 0) return IDASH_B200_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<CtView> vin(n), vout(n);
    int rc;
    bool one_launch = n > 1 && n <= RG_MAX_BATCHES;
    const idash_b200_model *widest = ms[0];
    for (uint32_t b = 0; b < n; ++b) {
        if (!ms[b]) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_multi_model: null model %u", b);
        if (ms[b]->device != c->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_multi_model: model %u lives on device %d, ctx on %d", b, ms[b]->device, c->device);
        if ((rc = make_view(&in[b], false, &vin[b], "cloud_eval_device_multi_model(in)"))) return rc;
        if ((rc = make_view(&out[b], true, &vout[b], "cloud_eval_device_multi_model(out)"))) return rc;
        const idash_b200_layout *L = ms[b]->layout, *L0 = ms[0]->layout;
        if (vout[b].count != L->n_rows) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_device_multi_model: out[%u].count != model rows", b);
        one_launch = one_launch && ring_selected(c, L) && L->n_overflow_rows == 0 && L->NR == L0->NR && L->RS == L0->RS && L->n_rows == L0->n_rows &&
                     L->tiles.size() == L0->tiles.size() &&
                     !vin[b].records && vin[b].index == nullptr && vin[b].count == vin[0].count && vin[b].stride == vin[0].stride &&
                     (vin[b].variance != nullptr) == (vin[0].variance != nullptr) &&
                     vout[b].records == vout[0].records && vout[b].stride == vout[0].stride &&
                     (vout[b].index != nullptr) == (vout[0].index != nullptr) && (vout[b].variance != nullptr) == (vout[0].variance != nullptr);
        if (L->tile_kmax > widest->layout->tile_kmax) widest = ms[b];
    }
    if (one_launch) {
        RingPlan rp;
        uint64_t blocks = 0;
        for (uint32_t b = 0; b < n; ++b) blocks = std::max<uint64_t>(blocks, (ms[b]->layout->tiles.back().f_base >> 5) + IDASH_B200_TILE_KMAX / 32u);
        one_launch = ring_plan(c, widest->layout, widest->layout->tiles.size(), n, &rp) && blocks < (1u << RG_BATCH_SHIFT);
    }
    // the widest model sizes the shared-memory plan; its arrays also stand in RingParams' single-model fields (unused by a batched launch)
    if (one_launch) return launch_cloud(c, widest, vin[0], vout[0], nullptr, (cudaStream_t) stream, nullptr, n, vin.data(), vout.data(), ms);
    for (uint32_t b = 0; b < n; ++b)
        if ((rc = launch_cloud(c, ms[b], vin[b], vout[b], nullptr, (cudaStream_t) stream))) return rc;
    return IDASH_B200_OK;
}

extern "C" int idash_b200_check_device_status(idash_b200_ctx *c) {
    clear_error();
    if (!c) return set_error(IDASH_B200_ERR_INVALID, "check_device_status: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int st = 0;
    CUDA_TRY(cudaMemcpy(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {
        CUDA_TRY(cudaMemset(c->d_status, 0, sizeof(int)));
        return set_error(IDASH_B200_ERR_MISSING_INPUT, "cloud_eval: the model references an input ciphertext that was not supplied");
    }
    return IDASH_B200_OK;
}

// Stages a host ciphertext array into ctx-owned device memory and returns its device view.
static int stage_in(idash_b200_ctx *c, const idash_b200_cts *a, DevBuf &buf, DevBuf &idx_buf, DevBuf &var_buf, CtView *v,
                    const char *what) {
    if (!a) return set_error(IDASH_B200_ERR_INVALID, "%s: null ciphertext array", what);
    if (a->count && !a->data) return set_error(IDASH_B200_ERR_INVALID, "%s: null data pointer", what);
    memset(v, 0, sizeof(*v));
    v->count = a->count;
    int rc;
    if (a->layout == IDASH_B200_LAYOUT_PACKED) {
        const size_t bytes = (size_t) a->count * IDASH_B200_CT_BYTES;
        if ((rc = buf.ensure(bytes + 16))) return rc;
        if (bytes) CUDA_TRY(cudaMemcpyAsync(buf.p, a->data, bytes, cudaMemcpyHostToDevice, c->stream));
        v->words = (uint8_t *) buf.p;
        v->stride = IDASH_B200_CT_BYTES;
        if (a->index) {
            if ((rc = idx_buf.ensure((size_t) a->count * 4 + 4))) return rc;
            if (a->count) CUDA_TRY(cudaMemcpyAsync(idx_buf.p, a->index, (size_t) a->count * 4, cudaMemcpyHostToDevice, c->stream));
            v->index = (uint32_t *) idx_buf.p;
        }
        if (a->variance) {
            if ((rc = var_buf.ensure((size_t) a->count * 8 + 8))) return rc;
            if (a->count) CUDA_TRY(cudaMemcpyAsync(var_buf.p, a->variance, (size_t) a->count * 8, cudaMemcpyHostToDevice, c->stream));
            v->variance = (double *) var_buf.p;
        }
    } else if (a->layout == IDASH_B200_LAYOUT_RECORDS) {
        if (a->index || a->variance) return set_error(IDASH_B200_ERR_INVALID, "%s: index/variance must be NULL for the RECORDS layout", what);
        const size_t bytes = (size_t) a->count * IDASH_B200_RECORD_BYTES;
        if ((rc = buf.ensure(bytes + 32))) return rc;
        // device image starts at +8 so that the word arrays are 16-byte aligned, as in the file
        if (bytes) CUDA_TRY(cudaMemcpyAsync((uint8_t *) buf.p + 8, a->data, bytes, cudaMemcpyHostToDevice, c->stream));
        v->words = (uint8_t *) buf.p + 16;
        v->stride = IDASH_B200_RECORD_BYTES;
        v->records = 1;
    } else {
        return set_error(IDASH_B200_ERR_INVALID, "%s: unknown layout %d", what, a->layout);
    }
    return IDASH_B200_OK;
}

// Caller rows are already in tile order: tile t = rows [64 t, 64 t + n_valid), so a tile range writes a contiguous slot range.
static bool rows_identity(const idash_b200_model *m) {
    if (m->rows_identity < 0) {
        const idash_b200_layout *L = m->layout;
        bool ok = !L->tiles.empty();
        for (size_t t = 0; ok && t < L->tiles.size(); ++t)
            for (uint32_t i = 0; ok && i < IDASH_B200_TILE_ROWS; ++i) {
                const uint32_t want = i < L->tiles[t].n_valid ? (uint32_t) (t * IDASH_B200_TILE_ROWS + i) : IDASH_B200_NO_ROW;
                ok = L->tile_rows[t * IDASH_B200_TILE_ROWS + i] == want;
            }
        m->rows_identity = ok ? 1 : 0;
    }
    return m->rows_identity == 1;
}

// cloud_eval_host for the common big case (ring kernel, caller rows in sorted order, outputs in row order): the target
// range is cut into pieces and the three engines run concurrently -- host->device copy of the input ciphertexts piece k+1
// needs (copy-in stream), kernels of piece k (compute stream), device->host copy of the outputs of piece k-1 (copy-out
// stream). PCIe is full duplex, so the input upload disappears behind the (5x larger) output download.
// [tile_lo, tile_hi): the tiles to evaluate (the whole model, or one GPU's contiguous target range of a multi-GPU evaluation);
// only their rows of `out` are written, only the input ciphertexts their bands touch are uploaded (PACKED identity inputs), and
// the device staging buffers are sized for that range.
static int cloud_eval_host_pipelined(idash_b200_ctx *c, const idash_b200_model *m, const idash_b200_cts *in, const idash_b200_cts *out,
                                     uint32_t tile_lo, uint32_t tile_hi) {
    const idash_b200_layout *L = m->layout;
    const uint32_t n_tiles = tile_hi - tile_lo;
    if (n_tiles == 0) return IDASH_B200_OK;
    // Pieces GROW (piece k holds a share ~ k + 1 of the tiles): the device->host copy -- the longest of the three streams -- can start as
    // soon as the first, small piece has been uploaded and evaluated, and every later piece is ready long before the copy of its
    // predecessor ends. (With 8 equal pieces the first download waited for 1/8 of the upload: ~1 ms of a 37 ms call.)
    const uint32_t P = std::max<uint32_t>(1u, std::min<uint32_t>(12u, n_tiles / 64u));
    const uint64_t P_tri = (uint64_t) P * (P + 1u) / 2u;
    const auto piece_begin = [&](uint32_t k) -> uint32_t { return tile_lo + (uint32_t) ((uint64_t) n_tiles * ((uint64_t) k * (k + 1u) / 2u) / P_tri); };
    if (!c->s_in) { CUDA_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)); CUDA_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking)); }
    while (c->ev_in.size() < P) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_in.push_back(a); c->ev_k.push_back(b);
    }
    int rc;
    CtView vin, vout;
    const uint64_t row_lo_all = (uint64_t) tile_lo * IDASH_B200_TILE_ROWS;
    const uint64_t row_hi_all = std::min<uint64_t>(L->n_rows, (uint64_t) tile_hi * IDASH_B200_TILE_ROWS);
    const bool chunked_in = in->layout == IDASH_B200_LAYOUT_PACKED && !in->index;     // slot i holds ciphertext i
    uint64_t ct_lo = 0;
    if (!chunked_in) {
        if ((rc = stage_in(c, in, c->in_buf, c->aux_in_idx, c->aux_in_var, &vin, "cloud_eval_host(in)"))) return rc;
    } else {
        if (in->count && !in->data) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host(in): null data pointer");
        // ciphertexts the band of [tile_lo, tile_hi) touches: [ct_lo, ct_hi)
        uint64_t f_lo = UINT64_MAX, f_hi = 0;
        for (uint32_t t = tile_lo; t < tile_hi; ++t) {
            f_lo = std::min<uint64_t>(f_lo, L->tiles[t].f_base);
            f_hi = std::max<uint64_t>(f_hi, (uint64_t) L->tiles[t].f_base + L->tiles[t].K);
        }
        ct_lo = std::min<uint64_t>(f_lo / L->NR, in->count);
        const uint64_t ct_hi = std::min<uint64_t>((f_hi + L->NR - 1) / L->NR, in->count);
        memset(&vin, 0, sizeof(vin));
        vin.count = in->count;
        if ((rc = c->in_buf.ensure((size_t) (ct_hi - ct_lo) * IDASH_B200_CT_BYTES + 16))) return rc;
        vin.words = (uint8_t *) c->in_buf.p - ct_lo * IDASH_B200_CT_BYTES;      // virtual base: slot ct_lo is the first one resident
        vin.stride = IDASH_B200_CT_BYTES;
        if (in->variance) {
            if ((rc = c->aux_in_var.ensure((size_t) in->count * 8 + 8))) return rc;
            if (in->count) CUDA_TRY(cudaMemcpyAsync(c->aux_in_var.p, in->variance, (size_t) in->count * 8, cudaMemcpyHostToDevice, c->stream));
            vin.variance = (double *) c->aux_in_var.p;
        }
    }
    memset(&vout, 0, sizeof(vout));
    vout.count = out->count;
    const bool rec = out->layout == IDASH_B200_LAYOUT_RECORDS;
    const size_t ostride = rec ? IDASH_B200_RECORD_BYTES : IDASH_B200_CT_BYTES;
    const uint64_t n_out_rows = row_hi_all - row_lo_all;
    // device output buffer for the rows of this range only (virtual base like the input's)
    if ((rc = c->out_buf.ensure((size_t) n_out_rows * ostride + 32))) return rc;
    uint8_t *const obase = (uint8_t *) c->out_buf.p - row_lo_all * ostride;
    if (rec) {
        if (out->index || out->variance) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host(out): index/variance must be NULL for the RECORDS layout");
        vout.words = obase + 16; vout.stride = IDASH_B200_RECORD_BYTES; vout.records = 1;
    } else {
        vout.words = obase; vout.stride = IDASH_B200_CT_BYTES;
        if (out->index) { if ((rc = c->aux_out_idx.ensure((size_t) n_out_rows * 4 + 4))) return rc; vout.index = (uint32_t *) c->aux_out_idx.p - row_lo_all; }
        if (out->variance) { if ((rc = c->aux_out_var.ensure((size_t) n_out_rows * 8 + 8))) return rc; vout.variance = (double *) c->aux_out_var.p - row_lo_all; }
    }
    uint64_t uploaded = ct_lo, need = 0;
    uint32_t t_scan = tile_lo;
    bool first_piece = true;
    for (uint32_t k = 0; k < P; ++k) {
        Piece pc;
        pc.tile_lo = piece_begin(k);
        pc.tile_hi = k + 1 == P ? tile_hi : piece_begin(k + 1);
        if (pc.tile_hi == pc.tile_lo) continue;
        pc.row_lo = (uint64_t) pc.tile_lo * IDASH_B200_TILE_ROWS;
        pc.row_hi = std::min<uint64_t>(L->n_rows, (uint64_t) pc.tile_hi * IDASH_B200_TILE_ROWS);
        pc.first = first_piece;
        first_piece = false;
        if (chunked_in) {
            for (; t_scan < pc.tile_hi; ++t_scan) need = std::max<uint64_t>(need, (uint64_t) L->tiles[t_scan].f_base + L->tiles[t_scan].K);   // features
            const uint64_t upto = std::min<uint64_t>((need + L->NR - 1) / L->NR, in->count);                                            // ciphertexts
            if (upto > uploaded) {
                CUDA_TRY(cudaMemcpyAsync(vin.words + uploaded * IDASH_B200_CT_BYTES, (const uint8_t *) in->data + uploaded * IDASH_B200_CT_BYTES,
                                         (upto - uploaded) * IDASH_B200_CT_BYTES, cudaMemcpyHostToDevice, c->s_in));
                uploaded = upto;
            }
            CUDA_TRY(cudaEventRecord(c->ev_in[k], c->s_in));
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
        }
        if ((rc = launch_cloud(c, m, vin, vout, nullptr, c->stream, &pc))) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_out); return rc; }
        CUDA_TRY(cudaEventRecord(c->ev_k[k], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
        const uint64_t nr = pc.row_hi - pc.row_lo;
        if (nr) {
            if (rec) {
                CUDA_TRY(cudaMemcpyAsync((uint8_t *) out->data + pc.row_lo * ostride, obase + 8 + pc.row_lo * ostride, nr * ostride, cudaMemcpyDeviceToHost, c->s_out));
            } else {
                CUDA_TRY(cudaMemcpyAsync((uint8_t *) out->data + pc.row_lo * ostride, obase + pc.row_lo * ostride, nr * ostride, cudaMemcpyDeviceToHost, c->s_out));
            }
        }
    }
    // the small per-row arrays go last, in one piece: callers often pass pageable memory for them, and a copy to pageable
    // memory blocks the host thread, which would stop the pieces above from being enqueued ahead of the GPU
    if (!rec && n_out_rows) {
        if (out->index) CUDA_TRY(cudaMemcpyAsync(out->index + row_lo_all, vout.index + row_lo_all, (size_t) n_out_rows * 4, cudaMemcpyDeviceToHost, c->s_out));
        if (out->variance) CUDA_TRY(cudaMemcpyAsync(out->variance + row_lo_all, vout.variance + row_lo_all, (size_t) n_out_rows * 8, cudaMemcpyDeviceToHost, c->s_out));
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->s_out));
    CUDA_TRY(cudaStreamSynchronize(c->s_in));
    if (*c->h_status) {
        CUDA_TRY(cudaMemset(c->d_status, 0, sizeof(int)));
        return set_error(IDASH_B200_ERR_MISSING_INPUT, "cloud_eval: the model references an input ciphertext that was not supplied");
    }
    return IDASH_B200_OK;
}

static bool host_pipeline_applies(const idash_b200_ctx *c, const idash_b200_model *m, const idash_b200_cts *in) {
    const idash_b200_layout *L = m->layout;
    return in->count < NO_SLOT && ring_selected(c, L) && L->n_overflow_rows == 0 && rows_identity(m);
}

// One GPU's share of an evaluation that is sharded by contiguous target ranges: rows [row_begin, row_end) of the
// model, row_begin a multiple of 64 (a tile boundary). `out` describes the WHOLE output array (out->count = model rows); only
// the rows of the range are written. Needs a model whose caller rows are sorted by output bigIndex and which the ring kernel takes.
extern "C" int idash_b200_cloud_eval_host_rows(idash_b200_ctx *c, const idash_b200_model *m, const idash_b200_cts *in,
                                               const idash_b200_cts *out, uint64_t row_begin, uint64_t row_end) {
    clear_error();
    if (!c || !m || !in || !out) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: null argument");
    if (m->device != c->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: model lives on device %d, ctx on %d", m->device, c->device);
    CUDA_TRY(cudaSetDevice(c->device));
    const idash_b200_layout *L = m->layout;
    if (out->count != L->n_rows) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: out->count (%llu) != model rows (%llu)",
                                                  (unsigned long long) out->count, (unsigned long long) L->n_rows);
    if (row_begin > row_end || row_end > L->n_rows || row_begin % IDASH_B200_TILE_ROWS || (row_end % IDASH_B200_TILE_ROWS && row_end != L->n_rows))
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: rows [%llu, %llu) are not a tile-aligned range of the model",
                         (unsigned long long) row_begin, (unsigned long long) row_end);
    if (row_begin == row_end) return IDASH_B200_OK;
    if (out->count && !out->data) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: null output data pointer");
    if (out->layout != IDASH_B200_LAYOUT_PACKED && out->layout != IDASH_B200_LAYOUT_RECORDS)
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows(out): unknown layout %d", out->layout);
    if (!host_pipeline_applies(c, m, in))
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host_rows: needs a model with rows sorted by output bigIndex that the persistent "
                                                 "ring kernel evaluates whole (no overflow rows)");
    return cloud_eval_host_pipelined(c, m, in, out, (uint32_t) (row_begin / IDASH_B200_TILE_ROWS),
                                     (uint32_t) ((row_end + IDASH_B200_TILE_ROWS - 1) / IDASH_B200_TILE_ROWS));
}

// One evaluation sharded over several GPUs of this process, the data resident on the FIRST of them ("GPU-resident data ->
// cudaMemcpyPeerAsync over NVLink"). GPU g takes a contiguous tile-aligned target range; it receives the slab of input ciphertexts
// the range reads by a peer copy and WRITES ITS ROWS STRAIGHT INTO THE OUTPUT ARRAY ON GPU 0 -- the ring kernel's epilogue stores (and
// the per-row finalize kernel's) go through the peer mapping, so the gather is fused into the kernel: no staging buffer and no
// second pass over the 8 KB rows. Where peer access is not available the rows are staged locally and copied back.
extern "C" int idash_b200_cloud_eval_multi_device(uint32_t n_gpus, idash_b200_ctx *const *cs, const idash_b200_model *const *ms,
                                                  const idash_b200_cts *in, const idash_b200_cts *out) {
    clear_error();
    if (!n_gpus || !cs || !ms || !in || !out) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: null argument");
    for (uint32_t g = 0; g < n_gpus; ++g) {
        if (!cs[g] || !ms[g]) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: null ctx / model %u", g);
        if (ms[g]->device != cs[g]->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: model %u lives on device %d, ctx on %d", g, ms[g]->device, cs[g]->device);
        if (ms[g]->layout != ms[0]->layout) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: model %u is not a clone of model 0", g);
        for (uint32_t h = 0; h < g; ++h)
            if (cs[h]->device == cs[g]->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: device %d listed twice", cs[g]->device);
    }
    const idash_b200_layout *L = ms[0]->layout;
    if (in->layout != IDASH_B200_LAYOUT_PACKED || in->index)
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: inputs must be PACKED in identity order (slot i = ciphertext i)");
    CtView vin0, vout0;
    int rc;
    if ((rc = make_view(in, false, &vin0, "cloud_eval_multi_device(in)"))) return rc;
    if ((rc = make_view(out, true, &vout0, "cloud_eval_multi_device(out)"))) return rc;
    if (out->count != L->n_rows) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: out->count (%llu) != model rows (%llu)",
                                                  (unsigned long long) out->count, (unsigned long long) L->n_rows);
    if (L->n_rows == 0) return IDASH_B200_OK;
    if (n_gpus == 1) {
        CUDA_TRY(cudaSetDevice(cs[0]->device));
        if ((rc = launch_cloud(cs[0], ms[0], vin0, vout0, nullptr, cs[0]->stream))) return rc;
        CUDA_TRY(cudaStreamSynchronize(cs[0]->stream));
        return idash_b200_check_device_status(cs[0]);
    }
    for (uint32_t g = 0; g < n_gpus; ++g)
        if (!ring_selected(cs[g], L) || L->n_overflow_rows || !rows_identity(ms[g]))
            return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_multi_device: needs a model with rows sorted by output bigIndex that the persistent "
                                                     "ring kernel evaluates whole (no overflow rows)");
    const uint64_t T = L->tiles.size();
    const int dev0 = cs[0]->device;
    const size_t ostride = vout0.stride;
    struct Staged { uint32_t g; uint64_t row_lo, row_hi; };
    std::vector<Staged> staged;
    for (uint32_t g = 0; g < n_gpus; ++g) {
        idash_b200_ctx *c = cs[g];
        Piece pc;
        pc.tile_lo = (uint32_t) (T * g / n_gpus); pc.tile_hi = (uint32_t) (T * (g + 1) / n_gpus);
        if (pc.tile_lo == pc.tile_hi) continue;
        pc.row_lo = (uint64_t) pc.tile_lo * IDASH_B200_TILE_ROWS;
        pc.row_hi = std::min<uint64_t>(L->n_rows, (uint64_t) pc.tile_hi * IDASH_B200_TILE_ROWS);
        pc.first = true;
        CUDA_TRY(cudaSetDevice(c->device));
        CtView vin = vin0, vout = vout0;
        if (g != 0) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, dev0));
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(dev0, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); can = 0; }
                else cudaGetLastError();
            }
            // the slab this range reads: [ct_lo, ct_hi)
            uint64_t f_lo = UINT64_MAX, f_hi = 0;
            for (uint32_t t = pc.tile_lo; t < pc.tile_hi; ++t) {
                f_lo = std::min<uint64_t>(f_lo, L->tiles[t].f_base);
                f_hi = std::max<uint64_t>(f_hi, (uint64_t) L->tiles[t].f_base + L->tiles[t].K);
            }
            const uint64_t ct_lo = std::min<uint64_t>(f_lo / L->NR, in->count), ct_hi = std::min<uint64_t>((f_hi + L->NR - 1) / L->NR, in->count);
            if ((rc = c->in_buf.ensure((size_t) (ct_hi - ct_lo) * IDASH_B200_CT_BYTES + 16))) return rc;
            if (ct_hi > ct_lo)
                CUDA_TRY(cudaMemcpyPeerAsync(c->in_buf.p, c->device, vin0.words + ct_lo * IDASH_B200_CT_BYTES, dev0, (size_t) (ct_hi - ct_lo) * IDASH_B200_CT_BYTES, c->stream));
            vin.words = (uint8_t *) c->in_buf.p - ct_lo * IDASH_B200_CT_BYTES;          // virtual base: slot ct_lo is the first one resident
            if (vin0.variance) {
                if ((rc = c->aux_in_var.ensure((size_t) in->count * 8 + 8))) return rc;
                CUDA_TRY(cudaMemcpyPeerAsync(c->aux_in_var.p, c->device, vin0.variance, dev0, (size_t) in->count * 8, c->stream));
                vin.variance = (double *) c->aux_in_var.p;
            }
            if (!can) {
                // no peer mapping: rows of the range into a local buffer (virtual base like the input's), copied back below
                const uint64_t nr = pc.row_hi - pc.row_lo;
                if ((rc = c->out_buf.ensure((size_t) nr * ostride + 32))) return rc;
                uint8_t *const obase = (uint8_t *) c->out_buf.p + (vout0.records ? 16 : 0) - pc.row_lo * ostride;
                vout.words = obase;
                if (vout0.index) { if ((rc = c->aux_out_idx.ensure((size_t) nr * 4 + 4))) return rc; vout.index = (uint32_t *) c->aux_out_idx.p - pc.row_lo; }
                if (vout0.variance) { if ((rc = c->aux_out_var.ensure((size_t) nr * 8 + 8))) return rc; vout.variance = (double *) c->aux_out_var.p - pc.row_lo; }
                staged.push_back({g, pc.row_lo, pc.row_hi});
            }
        }
        if ((rc = launch_cloud(c, ms[g], vin, vout, nullptr, c->stream, &pc))) {
            for (uint32_t h = 0; h <= g; ++h) { cudaSetDevice(cs[h]->device); cudaStreamSynchronize(cs[h]->stream); }
            return rc;
        }
    }
    for (const Staged &s : staged) {
        idash_b200_ctx *c = cs[s.g];
        CUDA_TRY(cudaSetDevice(c->device));
        const uint64_t nr = s.row_hi - s.row_lo;
        const size_t head = vout0.records ? 8 : 0;          // a record starts 8 bytes before its words
        CUDA_TRY(cudaMemcpyPeerAsync(vout0.words - head + s.row_lo * ostride, dev0, (uint8_t *) c->out_buf.p + (vout0.records ? 16 : 0) - head, c->device, nr * ostride, c->stream));
        if (vout0.index) CUDA_TRY(cudaMemcpyPeerAsync(vout0.index + s.row_lo, dev0, c->aux_out_idx.p, c->device, nr * 4, c->stream));
        if (vout0.variance) CUDA_TRY(cudaMemcpyPeerAsync(vout0.variance + s.row_lo, dev0, c->aux_out_var.p, c->device, nr * 8, c->stream));
    }
    int result = IDASH_B200_OK;
    for (uint32_t g = 0; g < n_gpus; ++g) {
        CUDA_TRY(cudaSetDevice(cs[g]->device));
        CUDA_TRY(cudaStreamSynchronize(cs[g]->stream));
        const int r = idash_b200_check_device_status(cs[g]);
        if (r != IDASH_B200_OK && result == IDASH_B200_OK) result = r;
    }
    return result;
}

extern "C" int idash_b200_model_input_range(const idash_b200_model *m, uint64_t row_begin, uint64_t row_end, uint32_t *ct_begin, uint32_t *ct_end) {
    clear_error();
    if (!m || !ct_begin || !ct_end) return set_error(IDASH_B200_ERR_INVALID, "model_input_range: null argument");
    const idash_b200_layout *L = m->layout;
    if (row_begin > row_end || row_end > L->n_rows || row_begin % IDASH_B200_TILE_ROWS || (row_end % IDASH_B200_TILE_ROWS && row_end != L->n_rows))
        return set_error(IDASH_B200_ERR_INVALID, "model_input_range: rows [%llu, %llu) are not a tile-aligned range of the model",
                         (unsigned long long) row_begin, (unsigned long long) row_end);
    if (L->tiles.empty() || L->n_overflow_rows || !rows_identity(m))
        return set_error(IDASH_B200_ERR_INVALID, "model_input_range: needs a model with rows sorted by output bigIndex and no overflow rows");
    uint64_t f_lo = UINT64_MAX, f_hi = 0;
    for (uint64_t t = row_begin / IDASH_B200_TILE_ROWS; t < (row_end + IDASH_B200_TILE_ROWS - 1) / IDASH_B200_TILE_ROWS; ++t) {
        f_lo = std::min<uint64_t>(f_lo, L->tiles[t].f_base);
        f_hi = std::max<uint64_t>(f_hi, (uint64_t) L->tiles[t].f_base + L->tiles[t].K);
    }
    if (f_lo > f_hi) { *ct_begin = *ct_end = 0; return IDASH_B200_OK; }
    *ct_begin = (uint32_t) (f_lo / L->NR);
    *ct_end = (uint32_t) ((f_hi + L->NR - 1) / L->NR);
    return IDASH_B200_OK;
}

extern "C" int idash_b200_cloud_eval_host(idash_b200_ctx *c, const idash_b200_model *m, const idash_b200_cts *in,
                                          const idash_b200_cts *out, const uint32_t *slot_of_row) {
    clear_error();
    if (!c || !m || !in || !out) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host: null argument");
    if (m->device != c->device) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host: model lives on device %d, ctx on %d", m->device, c->device);
    CUDA_TRY(cudaSetDevice(c->device));
    const idash_b200_layout *L = m->layout;
    if (out->count != L->n_rows) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host: out->count (%llu) != model rows (%llu)",
                                                  (unsigned long long) out->count, (unsigned long long) L->n_rows);
    if (out->count && !out->data) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host: null output data pointer");
    if (out->layout != IDASH_B200_LAYOUT_PACKED && out->layout != IDASH_B200_LAYOUT_RECORDS)
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host(out): unknown layout %d", out->layout);
    if (!slot_of_row && L->tiles.size() >= 128 && host_pipeline_applies(c, m, in) && !getenv("IDASH_B200_NO_PIPELINE"))
        return cloud_eval_host_pipelined(c, m, in, out, 0u, (uint32_t) L->tiles.size());
    int rc;
    CtView vin, vout;
    if ((rc = stage_in(c, in, c->in_buf, c->aux_in_idx, c->aux_in_var, &vin, "cloud_eval_host(in)"))) return rc;

    memset(&vout, 0, sizeof(vout));
    vout.count = out->count;
    const uint32_t *d_slot_of_row = nullptr;
    if (slot_of_row) {
        for (uint64_t r = 0; r < out->count; ++r)
            if (slot_of_row[r] >= out->count) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host: slot_of_row[%llu] out of range", (unsigned long long) r);
        if ((rc = c->row_slot_buf.ensure((size_t) out->count * 4 + 4))) return rc;
        if (out->count) CUDA_TRY(cudaMemcpyAsync(c->row_slot_buf.p, slot_of_row, (size_t) out->count * 4, cudaMemcpyHostToDevice, c->stream));
        d_slot_of_row = (const uint32_t *) c->row_slot_buf.p;
    }
    size_t out_bytes;
    if (out->layout == IDASH_B200_LAYOUT_PACKED) {
        out_bytes = (size_t) out->count * IDASH_B200_CT_BYTES;
        if ((rc = c->out_buf.ensure(out_bytes + 16))) return rc;
        vout.words = (uint8_t *) c->out_buf.p;
        vout.stride = IDASH_B200_CT_BYTES;
        if (out->index) { if ((rc = c->aux_out_idx.ensure((size_t) out->count * 4 + 4))) return rc; vout.index = (uint32_t *) c->aux_out_idx.p; }
        if (out->variance) { if ((rc = c->aux_out_var.ensure((size_t) out->count * 8 + 8))) return rc; vout.variance = (double *) c->aux_out_var.p; }
    } else if (out->layout == IDASH_B200_LAYOUT_RECORDS) {
        if (out->index || out->variance) return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host(out): index/variance must be NULL for the RECORDS layout");
        out_bytes = (size_t) out->count * IDASH_B200_RECORD_BYTES;
        if ((rc = c->out_buf.ensure(out_bytes + 32))) return rc;
        vout.words = (uint8_t *) c->out_buf.p + 16;
        vout.stride = IDASH_B200_RECORD_BYTES;
        vout.records = 1;
    } else {
        return set_error(IDASH_B200_ERR_INVALID, "cloud_eval_host(out): unknown layout %d", out->layout);
    }
    if ((rc = launch_cloud(c, m, vin, vout, d_slot_of_row, c->stream))) return rc;
    if (out->count) {
        if (vout.records) {
            CUDA_TRY(cudaMemcpyAsync(out->data, (uint8_t *) c->out_buf.p + 8, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        } else {
            CUDA_TRY(cudaMemcpyAsync(out->data, c->out_buf.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
            if (out->index) CUDA_TRY(cudaMemcpyAsync(out->index, vout.index, (size_t) out->count * 4, cudaMemcpyDeviceToHost, c->stream));
            if (out->variance) CUDA_TRY(cudaMemcpyAsync(out->variance, vout.variance, (size_t) out->count * 8, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (*c->h_status) {
        CUDA_TRY(cudaMemset(c->d_status, 0, sizeof(int)));
        return set_error(IDASH_B200_ERR_MISSING_INPUT, "cloud_eval: the model references an input ciphertext that was not supplied");
    }
    return IDASH_B200_OK;
}

static int pack_key(const int32_t *key, KeyBits *kb) {
    memset(kb, 0, sizeof(*kb));
    for (uint32_t i = 0; i < POLY_N; ++i) {
        if (key[i] != 0 && key[i] != 1) return set_error(IDASH_B200_ERR_INVALID, "decrypt: key coefficient %u is %d, expected a binary key", i, key[i]);
        if (key[i]) kb->w[i >> 5] |= 1u << (i & 31);
    }
    return IDASH_B200_OK;
}

// Tensor map (TMA descriptor) of the b polynomials of a ciphertext array: 1024 words x count rows, row pitch = the ciphertext stride
// (8192 packed, 8208 in a record stream -- both multiples of 16), box = 128 words x 16 ciphertexts. Encoded by the driver
// (cuTensorMapEncodeTiled, looked up through the runtime: the library does not link libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_b_map(const CtView &in, CUtensorMap *map) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return set_error(IDASH_B200_ERR_CUDA, "decrypt: the driver does not export cuTensorMapEncodeTiled");
        encode = (EncodeTiledFn) fn;
    }
    const cuuint64_t dims[2] = {POLY_N, (cuuint64_t) in.count}, pitch[1] = {(cuuint64_t) in.stride};
    const cuuint32_t box[2] = {128u, 16u}, unit[2] = {1u, 1u};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *) (in.words + 4u * POLY_N), dims, pitch, box, unit, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(IDASH_B200_ERR_CUDA, "decrypt: cuTensorMapEncodeTiled failed (%d)", (int) r);
    return IDASH_B200_OK;
}

static int launch_decrypt(idash_b200_ctx *c, const KeyBits &kb, uint32_t S, const CtView &in, float *d_scores, uint32_t *d_phase, cudaStream_t st) {
    if (in.count == 0) return IDASH_B200_OK;
    const bool timed = c->t_used < (int) c->t_begin.size();
    if (timed) CUDA_TRY(cudaEventRecord(c->t_begin[c->t_used], st));
    bool pair_launched = false;
    if (c->decrypt_choice == IDASH_B200_DECRYPT_TENSOR_PAIR || (c->decrypt_choice == IDASH_B200_DECRYPT_AUTO && c->sm_count >= 2)) {
        // K4p: CTA pairs (tcgen05.mma.cta_group::2), two groups of operand slots resident per CTA
        DecTcParams p;
        memset(&p, 0, sizeof(p));
        p.in = in;
        p.n_ct = in.count;
        p.n_groups = (in.count + DT_CTS - 1) / DT_CTS;
        p.S = S;
        p.n_slots = 16u;
        p.n_bstages = 3u;
        p.scores = d_scores;
        p.phase = d_phase;
        p.key = kb;
        uint64_t grid = std::min<uint64_t>(2 * p.n_groups, (uint64_t) c->sm_count) & ~1ull;
        if (const char *gs = debug_env("IDASH_B200_DECRYPT_GRID")) grid = std::max<uint64_t>(2, std::min<uint64_t>(grid, (uint64_t) atoi(gs)) & ~1ull);
        if (const char *ns = debug_env("IDASH_B200_DECRYPT_SLOTS")) p.n_slots = std::max<uint32_t>(DT_GROUP_SLOTS, std::min<uint32_t>(DP_MAX_SLOTS, (uint32_t) atoi(ns)));
        if (const char *bs = debug_env("IDASH_B200_DECRYPT_BSTAGES")) p.n_bstages = std::max<uint32_t>(2u, std::min<uint32_t>(DT_MAX_BSTAGES, (uint32_t) atoi(bs)));
        if (const char *ko = debug_env("IDASH_B200_DECRYPT_KNOCKOUT")) p.knockout = (uint32_t) atoi(ko);
        while (dec_pair_smem_bytes(p.n_slots, p.n_bstages) > RG_SMEM_MAX && p.n_slots > DT_GROUP_SLOTS) --p.n_slots;
        const size_t smem = dec_pair_smem_bytes(p.n_slots, p.n_bstages);
        CUtensorMap bmap;
        if (int rc = make_b_map(in, &bmap)) return rc;
        if (in.stride == IDASH_B200_RECORD_BYTES) {
            if (d_phase) decrypt_pair_kernel<IDASH_B200_RECORD_BYTES, true><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
            else decrypt_pair_kernel<IDASH_B200_RECORD_BYTES, false><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
        } else {
            if (d_phase) decrypt_pair_kernel<IDASH_B200_CT_BYTES, true><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
            else decrypt_pair_kernel<IDASH_B200_CT_BYTES, false><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
        }
        // a device that cannot place the pairs (a cluster launch needs both SMs of a TPC): AUTO falls back to one CTA per SM
        const cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) { pair_launched = true; c->last_decrypt_kernel = IDASH_B200_DECRYPT_TENSOR_PAIR; }
        else if (c->decrypt_choice == IDASH_B200_DECRYPT_TENSOR_PAIR) return set_error(IDASH_B200_ERR_CUDA, "decrypt: cluster launch failed: %s", cudaGetErrorString(e));
    }
    if (pair_launched) {
    } else if (c->decrypt_choice != IDASH_B200_DECRYPT_IADD) {
        DecTcParams p;
        memset(&p, 0, sizeof(p));
        p.in = in;
        p.n_ct = in.count;
        p.n_groups = (in.count + DT_CTS - 1) / DT_CTS;
        p.S = S;
        // shared memory: Toeplitz table 32 KB + operand ring (16.5 KB per slot) + b ring (8 KB per stage = the b words of a j block of
        // half a group). 10 slots = two slots of lookahead beyond a group; 3 b stages = one and a half j blocks in flight (measured with the
        // tensor-copy loader: 10 + 3 and 10 + 2 0.684 ms, 9 + 5 / 9 + 4 / 9 + 3 0.696 ms on the same box)
        p.n_slots = 10u;
        p.n_bstages = 3u;
        if (const char *bs = debug_env("IDASH_B200_DECRYPT_BSTAGES")) p.n_bstages = std::max<uint32_t>(2u, std::min<uint32_t>(DT_MAX_BSTAGES, (uint32_t) atoi(bs)));
        p.scores = d_scores;
        p.phase = d_phase;
        p.key = kb;
        uint64_t grid = std::min<uint64_t>(p.n_groups, (uint64_t) c->sm_count);
        if (const char *gs = debug_env("IDASH_B200_DECRYPT_GRID")) grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, (uint64_t) atoi(gs)));   // tests: many groups per CTA
        if (const char *ns = debug_env("IDASH_B200_DECRYPT_SLOTS")) p.n_slots = std::max<uint32_t>(DT_GROUP_SLOTS, std::min<uint32_t>(DT_MAX_SLOTS, (uint32_t) atoi(ns)));
        if (const char *ko = debug_env("IDASH_B200_DECRYPT_KNOCKOUT")) p.knockout = (uint32_t) atoi(ko);
        while (dec_tc_smem_bytes(p.n_slots, p.n_bstages) > RG_SMEM_MAX && p.n_slots > DT_GROUP_SLOTS) --p.n_slots;
        while (dec_tc_smem_bytes(p.n_slots, p.n_bstages) > RG_SMEM_MAX && p.n_bstages > 2u) --p.n_bstages;
        const size_t smem = dec_tc_smem_bytes(p.n_slots, p.n_bstages);
        CUtensorMap bmap;
        if (int rc = make_b_map(in, &bmap)) return rc;
        if (in.stride == IDASH_B200_RECORD_BYTES) {
            if (d_phase) decrypt_tc_kernel<IDASH_B200_RECORD_BYTES, true><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
            else decrypt_tc_kernel<IDASH_B200_RECORD_BYTES, false><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
        } else {
            if (d_phase) decrypt_tc_kernel<IDASH_B200_CT_BYTES, true><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
            else decrypt_tc_kernel<IDASH_B200_CT_BYTES, false><<<(unsigned) grid, DT_THREADS, smem, st>>>(p, bmap);
        }
        c->last_decrypt_kernel = IDASH_B200_DECRYPT_TENSOR;
    } else {
        const unsigned grid = (unsigned) std::min<uint64_t>(in.count, (uint64_t) c->sm_count * 16);
        decrypt_kernel<<<grid, 128, 0, st>>>(in, in.count, kb, S, d_scores, d_phase);
        c->last_decrypt_kernel = IDASH_B200_DECRYPT_IADD;
    }
    c->launches++;
    if (timed) CUDA_TRY(cudaEventRecord(c->t_end[c->t_used++], st));
    CUDA_TRY(cudaGetLastError());
    return IDASH_B200_OK;
}

extern "C" int idash_b200_decrypt_device(idash_b200_ctx *c, const int32_t *key_host, uint32_t S, const idash_b200_cts *in,
                                         float *scores, uint32_t *phase, void *stream) {
    clear_error();
    if (!c || !key_host || !in) return set_error(IDASH_B200_ERR_INVALID, "decrypt_device: null argument");
    if (S > POLY_N) return set_error(IDASH_B200_ERR_INVALID, "decrypt_device: num_samples %u > 1024", S);
    if (phase && ((uintptr_t) phase & 15u)) return set_error(IDASH_B200_ERR_INVALID, "decrypt_device: phase must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(c->device));
    KeyBits kb;
    int rc;
    if ((rc = pack_key(key_host, &kb))) return rc;
    CtView vin;
    if ((rc = make_view(in, false, &vin, "decrypt_device(in)"))) return rc;
    return launch_decrypt(c, kb, S, vin, scores, phase, (cudaStream_t) stream);
}

extern "C" int idash_b200_decrypt_host(idash_b200_ctx *c, const int32_t *key, uint32_t S, const idash_b200_cts *in,
                                       float *scores, uint32_t *phase) {
    clear_error();
    if (!c || !key || !in) return set_error(IDASH_B200_ERR_INVALID, "decrypt_host: null argument");
    if (S > POLY_N) return set_error(IDASH_B200_ERR_INVALID, "decrypt_host: num_samples %u > 1024", S);
    CUDA_TRY(cudaSetDevice(c->device));
    KeyBits kb;
    int rc;
    if ((rc = pack_key(key, &kb))) return rc;
    CtView vin;
    if ((rc = stage_in(c, in, c->in_buf, c->aux_in_idx, c->aux_in_var, &vin, "decrypt_host(in)"))) return rc;
    float *d_scores = nullptr;
    uint32_t *d_phase = nullptr;
    const size_t sbytes = (size_t) in->count * S * sizeof(float), pbytes = (size_t) in->count * POLY_N * 4;
    if (scores) { if ((rc = c->scores_buf.ensure(sbytes + 16))) return rc; d_scores = (float *) c->scores_buf.p; }
    if (phase) { if ((rc = c->phase_buf.ensure(pbytes + 16))) return rc; d_phase = (uint32_t *) c->phase_buf.p; }
    if ((rc = launch_decrypt(c, kb, S, vin, d_scores, d_phase, c->stream))) return rc;
    if (scores && sbytes) CUDA_TRY(cudaMemcpyAsync(scores, d_scores, sbytes, cudaMemcpyDeviceToHost, c->stream));
    if (phase && pbytes) CUDA_TRY(cudaMemcpyAsync(phase, d_phase, pbytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return IDASH_B200_OK;
}
