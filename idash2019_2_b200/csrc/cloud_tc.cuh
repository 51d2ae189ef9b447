// K2: cloud evaluation on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
// Included by idash_b200.cu after the device views (CtView, load_rotated, lookup_slot) are defined.
//
// One CTA = one band tile (64 consecutive output rows, include/idash_b200_layout.h) x one 128-word slice
// of the 2048-word ciphertext axis:
//     D[word][row] = sum_k X[f_base + k][word] * coef[row][k]        M = 128 words, N = 64 rows, K = band
// restating the coefficient loop of cloud_compute_score (eval/idash.cpp:800-819, tLweAddMulTo ->
// torusPolynomialAddMulZTo, toruspolynomial-functions.cpp:97-103) as a limb-split integer GEMM:
//     X = sum_j 2^(8j) X_j (u8), coef = c_lo + 256 c_hi (balanced s8 limbs)
//     P_w = X_w c_lo + X_(w-1) c_hi  (w = 0..3, int32 in TMEM),   out = sum_w 2^(8w) P_w  mod 2^32
// Products with weight 2^32 and above vanish mod 2^32. The coefficient operand is the 128-row matrix
// [c_lo | c_hi], so ONE N = 128 MMA of plane X_j adds X_j c_lo to P_j and X_j c_hi to P_(j+1) (adjacent TMEM
// columns): 3 such MMAs + one N = 64 MMA (X_3 c_lo) per 32 features give the exact Torus32 result.
// Phases of a CTA:
//   1. copy the tile's coefficient image (already in the K-major no-swizzle operand layout) to smem
//   2. stage A: every thread loads 16 consecutive (rotated, sign-corrected) words of one input
//      ciphertext, splits them into 4 byte planes (8 PRMT per 4 words) and stores one 16-byte chunk per
//      plane in the MN-major no-swizzle operand layout (core matrix = 8 features x 16 words)
//   3. allocate 256 TMEM columns (blocks while two other CTAs of the SM hold theirs, which overlaps this
//      CTA's staging with their epilogues), one thread issues the MMAs and commits to an mbarrier
//   4. epilogue: tcgen05.ld the 4 accumulators, recombine with shifts, add bias * 2^18 on b[0..S)
//      (idash.cpp:805-810), zero b[RS..N) (idash.cpp:839-841), streaming 128-byte-per-warp stores.
#pragma once

#define TC_TN IDASH_B200_TILE_ROWS   // 64 rows per tile
#ifndef TC_A_SBO
#define TC_A_SBO 144u                // bytes between 16-word groups of one 8-feature core-matrix row block (128 + 16 pad:
                                     // conflict-free 128-bit stores, verified by tools/umma_probe.cu)
#endif
#define TC_A_LBO (8u * TC_A_SBO)     // bytes between 8-feature groups
#define TC_B_SBO 128u                // 8 rows x 16 bytes
#define TC_B_LBO (2u * TC_TN * 16u)  // bytes between 16-feature halves of the [c_lo | c_hi] operand
#define TC_B_CHUNK (2u * 32u * TC_TN)  // coefficient image of one 32-feature K step: c_lo (2048 B) then c_hi
#define TC_THREADS 256

struct TcParams {
    const idash_b200_tile *tiles;
    const uint32_t *tile_rows;
    const int32_t *tile_bias;
    const uint8_t *tile_coef;
    const uint32_t *tile_used;
    uint32_t n_tiles;
    CtView in, out;
    const uint32_t *slot_of_ct;
    uint32_t n_ct_slots;
    const uint32_t *slot_of_row;
    uint32_t S, NR, RS;
    int *status;
};

__host__ __device__ constexpr uint32_t tc_smem_bytes(uint32_t kmax) { return 4u * (kmax / 8u) * TC_A_LBO + 2u * kmax * TC_TN; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

// tcgen05 shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr >> 4) & 0x3FFFu) | ((uint64_t) ((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t) 1 << 46);
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// instruction descriptor: D = s32, A = u8 MN-major, B = s8 K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t n) {
    return (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// The MMAs of one 32-feature K step: da = descriptor of limb plane 0 (planes are plane_units 16-byte units
// apart), db = descriptor of the coefficient chunk, d0 = TMEM column of P_0, first = overwrite the accumulators.
__device__ __forceinline__ void tc_mma_kstep(uint32_t d0, uint64_t da, uint32_t plane_units, uint64_t db, bool first) {
    const uint32_t acc = first ? 0u : 1u;
    tc_mma(d0 + 0 * TC_TN, da + 0 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc);   // P0, P1  = X0 [c_lo | c_hi]
    tc_mma(d0 + 2 * TC_TN, da + 2 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc);   // P2, P3  = X2 [c_lo | c_hi]
    tc_mma(d0 + 1 * TC_TN, da + 1 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), 1u);    // P1, P2 += X1 [c_lo | c_hi]
    tc_mma(d0 + 3 * TC_TN, da + 3 * (uint64_t) plane_units, db, tc_idesc(TC_TN), 1u);        // P3     += X3 c_lo
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void stg32_stream(void *p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 3) cloud_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint64_t row_ptr_s[TC_TN];    // address of word 0 of this slice in every tile row; 0 = no such row
    __shared__ uint32_t row_bias_s[TC_TN];   // Constant * 2^18

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tile_id = blockIdx.x >> 4, slice = blockIdx.x & 15u;
    const uint4 *tp = reinterpret_cast<const uint4 *>(p.tiles + tile_id);
    const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    const uint32_t f_base = t0.x, K = t0.y, used_off = t1.x;
    const uint64_t b_off = (uint64_t) t0.z | ((uint64_t) t0.w << 32);
    const uint32_t plane_bytes = (K >> 3) * TC_A_LBO;
    uint8_t *sA = smem;
    uint8_t *sB = smem + 4u * plane_bytes;
    const uint32_t w_slice = slice * 128u;               // first word of the slice inside a ciphertext
    const uint32_t poly_off = w_slice & POLY_N;           // 0: polynomial a, 1024: polynomial b
    const uint32_t i_slice = w_slice & (POLY_N - 1);
    const bool is_b = poly_off != 0;
    const bool all_masked = is_b && i_slice >= p.RS;      // whole slice lies in b[RS..N): zeros

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < TC_TN) {
        const uint32_t row = __ldg(p.tile_rows + (uint64_t) tile_id * TC_TN + tid);
        uint64_t ptr = 0;
        if (row != IDASH_B200_NO_ROW) {
            const uint32_t oslot = p.slot_of_row ? __ldg(p.slot_of_row + row) : row;
            ptr = (uint64_t) (p.out.words + (uint64_t) oslot * p.out.stride + 4u * w_slice);
        }
        row_ptr_s[tid] = ptr;
        row_bias_s[tid] = (uint32_t) __ldg(p.tile_bias + (uint64_t) tile_id * TC_TN + tid) * (uint32_t) IDASH_B200_ONE_IN_T32;
    }

    if (!all_masked) {
        // ---- 1. coefficient image -> smem (2 limbs x K x 64 bytes, already in operand layout)
        const uint4 *gB = reinterpret_cast<const uint4 *>(p.tile_coef + b_off);
        for (uint32_t i = tid; i < K * 8u; i += TC_THREADS) reinterpret_cast<uint4 *>(sB)[i] = __ldg(gB + i);

        // ---- 2. stage A: unit u = (feature k, 16-word group mg)
        for (uint32_t u = tid; u < K * 8u; u += TC_THREADS) {
            const uint32_t k = u >> 3, mg = u & 7u;
            const uint32_t f = f_base + k;
            uint32_t ct = f, shift = 0;
            if (MODE != 0) { ct = f / p.NR; shift = (f - ct * p.NR) * p.RS; }
            uint32_t slot = NO_SLOT;
            if (ct < p.n_ct_slots) slot = p.slot_of_ct ? __ldg(p.slot_of_ct + ct) : ct;
            uint4 w[4];
            if (slot == NO_SLOT) {
                // no such ciphertext: harmless when the band only pads over it, an error when a row uses it
                if ((__ldg(p.tile_used + used_off + (k >> 5)) >> (k & 31u)) & 1u) atomicOr(p.status, 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = make_uint4(0, 0, 0, 0);
            } else {
                const uint8_t *poly = p.in.words + (uint64_t) slot * p.in.stride + 4u * poly_off;
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = load_rotated<MODE>(poly, i_slice + mg * 16u + 4u * q, shift);
            }
            uint32_t limb[4][4];   // [plane j][q]: byte j of words 4q..4q+3
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t a0 = __byte_perm(w[q].x, w[q].y, 0x5140), a1 = __byte_perm(w[q].x, w[q].y, 0x7362);
                const uint32_t a2 = __byte_perm(w[q].z, w[q].w, 0x5140), a3 = __byte_perm(w[q].z, w[q].w, 0x7362);
                limb[0][q] = __byte_perm(a0, a2, 0x5410);
                limb[1][q] = __byte_perm(a0, a2, 0x7632);
                limb[2][q] = __byte_perm(a1, a3, 0x5410);
                limb[3][q] = __byte_perm(a1, a3, 0x7632);
            }
            uint8_t *dst = sA + (k >> 3) * TC_A_LBO + mg * TC_A_SBO + (k & 7u) * 16u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4 *>(dst + j * plane_bytes) = make_uint4(limb[j][0], limb[j][1], limb[j][2], limb[j][3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core

        // ---- 3. TMEM + MMA
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(4u * TC_TN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    uint32_t tmem = 0;
    if (!all_masked) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = tmem_base_s;
        if (tid == 0) {
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            const uint64_t da = tc_desc(a0, TC_A_LBO, TC_A_SBO), db = tc_desc(b0, TC_B_LBO, TC_B_SBO);
            for (uint32_t ks = 0; ks < (K >> 5); ++ks)
                tc_mma_kstep(tmem, da + (uint64_t) ((ks * 4u * TC_A_LBO) >> 4), plane_bytes >> 4, db + (uint64_t) ((ks * TC_B_CHUNK) >> 4), ks == 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mma_bar)), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- 4. epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 (words) and columns 32 (w / 4) .. +31 (rows)
    const uint32_t q4 = warp & 3u, half = warp >> 2;
    const uint32_t word_in_slice = q4 * 32u + lane;
    const uint32_t i_word = i_slice + word_in_slice;
    const bool bias_on = is_b && i_word < p.S;
    const bool zero_out = is_b && i_word >= p.RS;
#pragma unroll 1
    for (uint32_t chunk = 0; chunk < 4; ++chunk) {
        const uint32_t col0 = half * 32u + chunk * 8u;
        uint32_t v0[8], v1[8], v2[8], v3[8];
        if (!all_masked) {
            const uint32_t taddr = tmem + ((q4 * 32u) << 16) + col0;
            tc_ld8(taddr, v0);
            tc_ld8(taddr + TC_TN, v1);
            tc_ld8(taddr + 2 * TC_TN, v2);
            tc_ld8(taddr + 3 * TC_TN, v3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint64_t ptr = row_ptr_s[col0 + c];
            if (ptr == 0) continue;
            uint32_t v = 0;
            if (!all_masked) {
                v = v0[c] + (v1[c] << 8) + (v2[c] << 16) + (v3[c] << 24);
                if (bias_on) v += row_bias_s[col0 + c];
                if (zero_out) v = 0;
            }
            stg32_stream(reinterpret_cast<uint8_t *>(ptr) + 4u * word_in_slice, v);
        }
    }
    if (!all_masked) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(4u * TC_TN));
    }
}
