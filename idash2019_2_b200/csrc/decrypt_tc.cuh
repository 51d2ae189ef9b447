// K4t: decrypt_predictions (eval/idash.cpp:681-761) on the 5th-generation tensor cores.
//
// The reference computes phase = b - s * a in Z[X]/(X^1024 + 1) with a double-precision FFT
// (tLwePhase, tlwe-functions.cpp:64-71 -> torusPolynomialSubMulRFFT, polynomials.cpp:72-83); the exact product
// (torusPolynomialMultNaive, multiplication.cpp:53-65) is the contract here. With a binary key it is
//     prod[j] = sum_k T[j][k] a[k],   T[j][k] = s[j-k] (j >= k),  -s[1024+j-k] (j < k)     (T in {-1, 0, 1})
// i.e. one 1024 x 1024 negacyclic Toeplitz matrix applied to every ciphertext: a GEMM whose left operand is the
// same for all 242 646 ciphertexts. The CUDA-core kernel (decrypt_kernel) needs ~512 integer adds per output
// word and is INT32-issue bound (17 ms at iDASH scale); here the adds run as tcgen05.mma.kind::i8:
//     a = sum_l 2^(8l) a_l (unsigned byte planes),  P_l = T a_l  (|P_l| <= 1024 * 255, exact in int32 TMEM),
//     prod = sum_l 2^(8l) P_l mod 2^32,  phase = b - prod,  score = float(double(int32 phase) / 2^32)
//
// GEMM shape per MMA: D[j][n] += A[j][k'] B[k'][n],  M = 128 phase coefficients j (TMEM lanes),
// N = 128 = 4 byte planes x 32 ciphertexts (n = 32 l + ct), K = 32.
//   * A = T, signed bytes, MN-major, no swizzle. The coefficient axis of a is walked BACKWARDS (k' = 1023 - k),
//     so that A[j][k'] = t(j + k' - 1023) depends on j + k' only, with t(x) = s[x] (x >= 0), -s[x + 1024] (x < 0).
//     A core matrix (8 k' rows x 16 j bytes) at (j0, k0') is then a function of d = j0 + k0' (a multiple of 8):
//     254 distinct core matrices, 32.5 KB, built ONCE per CTA from the key bits. Every A tile of the 8 x 32
//     (j block, K step) grid is a descriptor into that table: start = 128 (16 jb + 4 ks), LBO (next 8 k') = 128 B,
//     SBO (next 16 j) = 256 B -- overlapping core matrices, read-only.
//   * B = byte planes of a, unsigned, K-major: rows of 16 consecutive k' of one (plane, ciphertext). Producer
//     threads load 64 contiguous bytes (16 coefficients) of one ciphertext, transpose bytes with PRMT (reversed
//     order) and store one 16-byte row per plane; lanes = ciphertexts, so a quarter-warp store is 128 contiguous
//     bytes (no bank conflicts). The operand lives in a shared-memory ring of 16 KB slots (128 k' x 128 n).
//   * TMEM: 2 stages x 2 j blocks x 128 columns = all 512 columns. A group of 32 ciphertexts takes 4 passes
//     (2 j blocks each) over its 8 ring slots: 256 MMAs ~ 16 k cycles, the same order as the HBM time of the
//     group's 384 KB (a, b in; scores out), so the kernel sits near both rooflines; measured numbers in DESIGN.md.
//   * roles (544 threads, one persistent CTA per SM, groups strided over the grid): warps 0-7 epilogue (lane
//     quadrant = warp % 4, j block of the pass = warp / 4; b is prefetched before the accumulator is ready),
//     warp 8 MMA issuer + TMEM owner, warps 9-16 producers (two 64-byte loads in flight per thread).
#pragma once

#define DT_CTS 32u                        // ciphertexts per group
#define DT_N 128u                         // MMA N = 4 planes x DT_CTS
#define DT_SLOT_BYTES 16384u              // one ring slot: 128 k' x 128 n bytes = 4 K steps
#define DT_GROUP_SLOTS 8u                 // 1024 k' per group
#define DT_B_LBO (16u * DT_N)             // bytes between 16-k' column blocks of a slot
#define DT_B_SBO 128u                     // 8 n rows x 16 bytes
#define DT_A_LBO 128u
#define DT_A_SBO 256u
#define DT_TOEP_CORES 254u
#define DT_TOEP_BYTES 32768u              // 254 x 128 = 32512, rounded up
#define DT_MAX_SLOTS 12u
#define DT_WARP_MMA 8u
#define DT_WARP_PROD 9u
#define DT_PROD_WARPS 8u
#define DT_THREADS ((DT_WARP_PROD + DT_PROD_WARPS) * 32u)

struct DecTcParams {
    CtView in;
    uint64_t n_ct;
    uint64_t n_groups;
    uint32_t S;
    uint32_t n_slots;       // ring slots (>= DT_GROUP_SLOTS)
    float *scores;          // [n_ct][S] or null
    uint32_t *phase;        // [n_ct][1024] or null
    KeyBits key;
};

__host__ __device__ constexpr uint32_t dec_tc_smem_bytes(uint32_t n_slots) { return DT_TOEP_BYTES + n_slots * DT_SLOT_BYTES; }

// instruction descriptor: D = s32, A = s8 MN-major, B = u8 K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t dec_idesc(uint32_t n) {
    return (2u << 4) | (1u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// byte plane l of four words, in REVERSED word order: result bytes = (w.w, w.z, w.y, w.x)[byte l]
__device__ __forceinline__ void dec_split_rev(const uint4 w, uint32_t &l0, uint32_t &l1, uint32_t &l2, uint32_t &l3) {
    const uint32_t wz01 = __byte_perm(w.w, w.z, 0x5140), wz23 = __byte_perm(w.w, w.z, 0x7362);   // (w0 z0 w1 z1), (w2 z2 w3 z3)
    const uint32_t yx01 = __byte_perm(w.y, w.x, 0x5140), yx23 = __byte_perm(w.y, w.x, 0x7362);   // (y0 x0 y1 x1), (y2 x2 y3 x3)
    l0 = __byte_perm(wz01, yx01, 0x5410);
    l1 = __byte_perm(wz01, yx01, 0x7632);
    l2 = __byte_perm(wz23, yx23, 0x5410);
    l3 = __byte_perm(wz23, yx23, 0x7632);
}

__device__ __forceinline__ uint32_t ldg32_nc(const void *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }

__global__ void __launch_bounds__(DT_THREADS, 1) decrypt_tc_kernel(const DecTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[DT_MAX_SLOTS], empty_bar[DT_MAX_SLOTS], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t key_s[32];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t NS = p.n_slots;
    uint8_t *toep = smem;
    uint8_t *ring = smem + DT_TOEP_BYTES;

    if (tid < 32) key_s[tid] = p.key.w[tid];
    if (tid == 0) {
        for (uint32_t i = 0; i < NS; ++i) { mbar_init(&full_bar[i], DT_PROD_WARPS * 32u); mbar_init(&empty_bar[i], 1); }
        for (uint32_t i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8u * 32u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == DT_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();
    // Toeplitz table: core matrix i (d = 8 i), row kk (k' offset), byte mm (j offset) = t(8 i + kk + mm - 1023)
    for (uint32_t idx = tid * 4u; idx < DT_TOEP_CORES * 128u; idx += DT_THREADS * 4u) {
        const int32_t x0 = (int32_t) (8u * (idx >> 7) + ((idx >> 4) & 7u) + (idx & 15u)) - 1023;
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int32_t x = x0 + b;
            const uint32_t xi = (uint32_t) (x >= 0 ? x : x + 1024);
            const uint32_t bit = (key_s[xi >> 5] >> (xi & 31u)) & 1u;
            const uint32_t v = x >= 0 ? bit : (0u - bit) & 0xFFu;
            word |= v << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(toep + idx) = word;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp < 8u) {
        // ---------------- epilogue: phase = b - sum_l 2^(8l) P_l, decode, store
        const uint32_t qd = warp & 3u, jl = warp >> 2;
        uint32_t pc = 0;
        for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
            const uint64_t ct0 = g * DT_CTS;
            const uint32_t n_here = (uint32_t) min((uint64_t) DT_CTS, p.n_ct - ct0);
            const uint8_t *b0 = p.in.words + ct0 * p.in.stride + 4u * POLY_N;
#pragma unroll 1
            for (uint32_t pass = 0; pass < 4; ++pass, ++pc) {
                const uint32_t stage = pc & 1u;
                const uint32_t j = (pass * 2u + jl) * 128u + qd * 32u + lane;
                uint32_t bw[DT_CTS];
#pragma unroll
                for (uint32_t c = 0; c < DT_CTS; ++c) bw[c] = c < n_here ? ldg32_nc(b0 + c * p.in.stride + 4u * j) : 0u;
                mbar_wait(&tfull_bar[stage], (pc >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem + ((qd * 32u) << 16) + stage * 256u + jl * DT_N;
                float *sc = p.scores ? p.scores + ct0 * p.S + j : nullptr;
                uint32_t *ph = p.phase ? p.phase + ct0 * POLY_N + j : nullptr;
#pragma unroll
                for (uint32_t chunk = 0; chunk < 4; ++chunk) {
                    uint32_t v0[8], v1[8], v2[8], v3[8];
                    tc_ld8(taddr + 0 * DT_CTS + chunk * 8u, v0);
                    tc_ld8(taddr + 1 * DT_CTS + chunk * 8u, v1);
                    tc_ld8(taddr + 2 * DT_CTS + chunk * 8u, v2);
                    tc_ld8(taddr + 3 * DT_CTS + chunk * 8u, v3);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (uint32_t c = 0; c < 8; ++c) {
                        const uint32_t cc = chunk * 8u + c;
                        if (cc < n_here) {
                            const uint32_t prod = v0[c] + (v1[c] << 8) + (v2[c] << 16) + (v3[c] << 24);
                            const uint32_t phs = bw[cc] - prod;
                            if (ph) stg32_stream(ph + (uint64_t) cc * POLY_N, phs);
                            // (float) (double(int32) / 2^32): one rounding to 24 bits, then an exact power-of-two scale --
                            // identical to idash.cpp:718 + numeric-functions.cpp:36-38
                            if (sc && j < p.S) stg32_stream(sc + (uint64_t) cc * p.S, __float_as_uint(__int2float_rn((int32_t) phs) * 2.3283064365386963e-10f));
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&tempty_bar[stage]);
            }
        }
    } else if (warp == DT_WARP_MMA) {
        // ---------------- MMA issuer
        const uint32_t leader = elect_one();
        const uint32_t toep_a = smem_u32(toep), ring_a = smem_u32(ring);
        uint32_t slot0 = 0, ph0 = 0, pc = 0;
        for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
#pragma unroll 1
            for (uint32_t pass = 0; pass < 4; ++pass, ++pc) {
                const uint32_t stage = pc & 1u;
                mbar_wait(&tempty_bar[stage], ((pc >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t slot = slot0, ph = ph0;
#pragma unroll 1
                for (uint32_t s = 0; s < DT_GROUP_SLOTS; ++s) {
                    if (pass == 0) {
                        mbar_wait(&full_bar[slot], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t b_a = ring_a + slot * DT_SLOT_BYTES;
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) {
                        const uint32_t ks = s * 4u + kk;
                        const uint64_t db = tc_desc(b_a + kk * 2u * DT_B_LBO, DT_B_LBO, DT_B_SBO);
#pragma unroll
                        for (uint32_t q = 0; q < 2; ++q) {
                            const uint32_t jb = pass * 2u + q;
                            const uint64_t da = tc_desc(toep_a + 128u * (16u * jb + 4u * ks), DT_A_LBO, DT_A_SBO);
                            tc_mma_p(tmem + stage * 256u + q * DT_N, da, db, dec_idesc(DT_N), ks != 0u, leader);
                        }
                    }
                    if (pass == 3 && leader) tc_commit(&empty_bar[slot]);   // the group is done with this slot
                    if (++slot == NS) { slot = 0; ph ^= 1u; }
                }
                if (leader) tc_commit(&tfull_bar[stage]);
                __syncwarp();
            }
            slot0 += DT_GROUP_SLOTS;
            if (slot0 >= NS) { slot0 -= NS; ph0 ^= 1u; }
        }
    } else {
        // ---------------- producers: unit i = (group, slot s): thread (ct, ub) stages k' block 8 s + ub of its ciphertext
        const uint32_t pt = tid - DT_WARP_PROD * 32u;
        const uint32_t ct_l = pt & 31u, ub = pt >> 5;
        const uint64_t my_groups = p.n_groups > blockIdx.x ? (p.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint64_t total = my_groups * DT_GROUP_SLOTS;
        auto load_unit = [&](uint64_t i, uint4 (&w)[4]) {
            if (i >= total) return;
            const uint64_t g = blockIdx.x + (i >> 3) * gridDim.x;
            const uint32_t kb = (uint32_t) (i & 7u) * 8u + ub;
            uint64_t ct = g * DT_CTS + ct_l;
            if (ct >= p.n_ct) ct = p.n_ct - 1;     // tail group: a valid address; the epilogue ignores these columns
            const uint8_t *src = p.in.words + ct * p.in.stride + 64u * (63u - kb);
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = ldg128(src + 16 * q);
        };
        uint32_t slot = 0, ph = 0;
        auto store_unit = [&](const uint4 (&w)[4]) {
            uint32_t l[4][4];   // [plane][16-byte row word]: row bytes 0..15 = coefficients k_start+15 .. k_start
#pragma unroll
            for (int q = 0; q < 4; ++q) dec_split_rev(w[3 - q], l[0][q], l[1][q], l[2][q], l[3][q]);
            mbar_wait(&empty_bar[slot], ph ^ 1u);
            uint8_t *dst = ring + slot * DT_SLOT_BYTES + ub * DT_B_LBO + ct_l * 16u;
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
                *reinterpret_cast<uint4 *>(dst + pl * (DT_CTS * 16u)) = make_uint4(l[pl][0], l[pl][1], l[pl][2], l[pl][3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_bar[slot]);
            if (++slot == NS) { slot = 0; ph ^= 1u; }
        };
        uint4 wa[4], wb[4];
        load_unit(0, wa);
        load_unit(1, wb);
        for (uint64_t i = 0; i < total; i += 2) {
            store_unit(wa);
            load_unit(i + 2, wa);
            if (i + 1 < total) {
                store_unit(wb);
                load_unit(i + 3, wb);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == DT_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
