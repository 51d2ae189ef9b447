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
// N = 128 = 32 ciphertexts x 4 byte planes (n = 4 ct + l), K = 32.
//   * A = T, signed bytes, MN-major, no swizzle. The coefficient axis of a is walked BACKWARDS (k' = 1023 - k),
//     so that A[j][k'] = t(j + k' - 1023) depends on j + k' only, with t(x) = s[x] (x >= 0), -s[x + 1024] (x < 0).
//     A core matrix (8 k' rows x 16 j bytes) at (j0, k0') is then a function of d = j0 + k0' (a multiple of 8):
//     254 distinct core matrices, 32.5 KB, built ONCE per CTA from the key bits. Every A tile of the 8 x 32
//     (j block, K step) grid is a descriptor into that table: start = 128 (16 jb + 4 ks), LBO (next 8 k') = 128 B,
//     SBO (next 16 j) = 256 B -- overlapping core matrices, read-only.
//   * B = byte planes of a, unsigned, K-major: rows of 16 consecutive k' of one (ciphertext, plane). A producer warp
//     loads 512 contiguous bytes of one ciphertext per instruction (fully coalesced: the first version had lanes =
//     ciphertexts, 32 cache lines per load instruction, and the LSU/L1 path -- not HBM -- bound the kernel); the 4
//     lanes that hold one 16-coefficient block split their words into byte planes (PRMT, reversed order), transpose
//     4 x 4 with 4 shuffles, and each stores one 16-byte row. The operand lives in a shared-memory ring of slots
//     (128 k' x 128 n, 16.5 KB with the bank-conflict pad).
//   * TMEM: 4 stages x 128 columns = all 512 columns, one j block each. A group of 32 ciphertexts takes 4 passes (PAIRS of j
//     blocks: tcgen05.mma.ws keeps a K step's byte-plane chunk in the collector buffer for both j blocks, 6 instead of 8 KB of
//     operand reads per MMA) over its 8 ring slots: 256 MMAs ~ 16 k cycles, the same order as the HBM time of the group's 384 KB
//     (a, b in; scores out), so the kernel sits near both rooflines; measured numbers in DESIGN.md 3.4.
//   * roles (448 threads, 128 registers per thread; one persistent CTA per SM, groups strided over the grid):
//     warps 0-3 epilogue (warp = TMEM lane quadrant; TMEM loads double-buffered; the b words come from a shared-memory ring),
//     warp 4 MMA issuer + TMEM owner, warps 5-12 producers (six 4 x 16-byte units in flight per thread; a unit is split and
//     transposed two units ahead of its store, so a freed slot is refilled by a wait and four stores), warp 13 b loader
//     (one thread: a tensor copy -- cp.async.bulk.tensor.2d -- of 128 words x 16 ciphertexts per half stage).
//     The first version had 8 epilogue warps (544 threads, 96 registers): they idled 78 % of the time while the register
//     cap limited the producers to 4 units.
//   * decrypt_pair.cuh is this kernel on CTA pairs (cta_group::2) and the default; this one remains selectable
//     (IDASH_B200_DECRYPT_TENSOR) and is what AUTO falls back to where a cluster launch is not possible.
#pragma once

#define DT_CTS 32u                        // ciphertexts per group
#define DT_N 128u                         // MMA N = 4 planes x DT_CTS
#define DT_GROUP_SLOTS 8u                 // 1024 k' per group
#define DT_B_LBO (16u * DT_N + 64u)       // bytes between 16-k' column blocks of a slot; the 64-byte pad makes the producers'
                                          // 128-bit stores (two k' blocks x four planes per quarter-warp) conflict-free
#define DT_B_SBO 128u                     // 8 n rows x 16 bytes
#define DT_SLOT_BYTES (8u * DT_B_LBO)     // one ring slot: 128 k' x 128 n bytes = 4 K steps (16896 with the pads)
#define DT_A_LBO 128u
#define DT_A_SBO 256u
#define DT_TOEP_CORES 254u
#define DT_TOEP_BYTES 32768u              // 254 x 128 = 32512, rounded up
#define DT_MAX_SLOTS 11u
#define DT_EPI_WARPS 4u
#define DT_WARP_MMA 4u
#define DT_WARP_PROD 5u
#define DT_PROD_WARPS 8u
#define DT_WARP_BLOAD (DT_WARP_PROD + DT_PROD_WARPS)          // one thread: bulk copies of the b words into the b ring
#define DT_THREADS ((DT_WARP_BLOAD + 1u) * 32u)
#define DT_B_STAGE_BYTES (16u * 512u)                          // one stage of the b ring: the 128 b words of a j block of HALF a group (16 ciphertexts)
#define DT_MAX_BSTAGES 8u

struct DecTcParams {
    CtView in;
    uint64_t n_ct;
    uint64_t n_groups;
    uint32_t S;
    uint32_t n_slots;       // ring slots (>= DT_GROUP_SLOTS)
    uint32_t n_bstages;     // stages of the b ring (>= 2; a j block of a group takes two)
    float *scores;          // [n_ct][S] or null
    uint32_t *phase;        // [n_ct][1024] or null
    KeyBits key;
    uint32_t knockout;      // profiling aid (IDASH_B200_DECRYPT_KNOCKOUT, results are wrong when non-zero): 1 no MMAs, 2 no output
                            // stores, 4 no b loads, 8 no a loads, 16 no operand stores, 32 no epilogue TMEM loads, 64 MMAs one j block
                            // at a time without the weight-stationary pairing
};

__host__ __device__ constexpr uint32_t dec_tc_smem_bytes(uint32_t n_slots, uint32_t n_bstages) {
    return DT_TOEP_BYTES + n_slots * DT_SLOT_BYTES + n_bstages * DT_B_STAGE_BYTES;
}

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

// One pass of one epilogue warp: 32 phase coefficients (lanes) x 32 ciphertexts. PHASE and FULL (all 32 ciphertexts of the group
// exist) are compile-time, so that every phase store is base + immediate and the score address is one pointer walk: ~9 instructions
// per output (the first version spent ~40 and the epilogue warps, not the tensor pipe or HBM, set the pace).
// The b words come from the b ring in shared memory (b_addr = this lane's word of ciphertext 0 of the stage; + 512 bytes per
// ciphertext), filled by bulk copies one or more j blocks ahead. They used to be 32-bit global loads prefetched one pass ahead into
// 32 registers per lane: 4 KB in flight per warp, 16 KB per SM -- with ~1 us of latency that caps the b stream at ~2.4 TB/s over the
// GPU, and the epilogue alone (b in, scores out) took 0.52 of the kernel's 0.71 ms (knock-outs, profiles/r02_decrypt_knockout.txt).
template <bool PHASE, bool FULL>
__device__ __forceinline__ void dec_epilogue_pass(uint32_t b_addr0, uint32_t b_addr1, uint64_t *bfull0, uint64_t *bfull1, uint32_t bpar0, uint32_t bpar1,
                                                  uint64_t *bempty0, uint64_t *bempty1, uint32_t n_here, uint32_t taddr, float *sc, uint32_t S,
                                                  bool sc_on, uint32_t *ph, uint64_t *tempty) {
    uint8_t *scp = reinterpret_cast<uint8_t *>(sc);   // walks down the 32 score rows of the group: + 4 S bytes per ciphertext
    const uint32_t s4 = 4u * S;
    // The accumulator is read in four chunks of 8 ciphertexts; the TMEM loads of chunk c + 1 are issued before chunk c is recombined
    // and stored (two register sets), so only the first round trip of a pass is exposed -- with one set the warp sat out four per pass,
    // ~1000 of the ~2500 cycles a pass took with the epilogue running alone (knock-outs, profiles/r02_decrypt_knockout.txt)
    uint32_t vv[2][4][8];     // vv[set][m][4 e + l]: plane l of ciphertext 8 chunk + 2 m + e
#pragma unroll
    for (uint32_t m = 0; m < 4; ++m) tc_ld8(taddr + m * 8u, vv[0][m]);
#pragma unroll
    for (uint32_t chunk = 0; chunk < 4; ++chunk) {
        uint32_t (&v)[4][8] = vv[chunk & 1u];
        uint32_t bw[8];
        // b words: ciphertexts 0-15 of the group from the first half stage, 16-31 from the second
        if (chunk == 0) mbar_wait(bfull0, bpar0);
        if (chunk == 2) mbar_wait(bfull1, bpar1);
        const uint32_t ba = (chunk < 2 ? b_addr0 : b_addr1) + (chunk & 1u) * 8u * 512u;
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bw[c]) : "r"(ba + c * 512u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (chunk < 3) {
#pragma unroll
            for (uint32_t m = 0; m < 4; ++m) tc_ld8(taddr + (chunk + 1u) * 32u + m * 8u, vv[(chunk + 1u) & 1u][m]);
        }
        if (chunk & 1u) {      // the half stage's words are in registers: hand it back to the loader
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive(chunk == 1 ? bempty0 : bempty1);
        }
        if (chunk == 3) {
            // the accumulator is in registers: hand the TMEM stage back before the last 8 outputs are computed and stored
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty);
        }
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const uint32_t cc = chunk * 8u + c;
            const uint32_t *q = &v[c >> 1][4u * (c & 1u)];
            const uint32_t phs = bw[c] - (q[0] + (q[1] << 8) + (q[2] << 16) + (q[3] << 24));
            const bool here = FULL || cc < n_here;
            if (PHASE && here) stg32_stream(ph + cc * POLY_N, phs);
            // (float) (double(int32) / 2^32): one rounding to 24 bits, then an exact power-of-two scale --
            // identical to idash.cpp:718 + numeric-functions.cpp:36-38. Only the store is predicated (j < S).
            const uint32_t fbits = __float_as_uint(__int2float_rn((int32_t) phs) * 2.3283064365386963e-10f);
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.u32 [%0], %1;\n\t}\n" ::"l"(scp), "r"(fbits), "r"((uint32_t) (sc_on && here)) : "memory");
            scp += s4;
        }
    }
}

template <uint32_t STRIDE, bool PHASE>
__global__ void __launch_bounds__(DT_THREADS, 1) decrypt_tc_kernel(const DecTcParams p, const __grid_constant__ CUtensorMap bmap) {   // 13 warps: 4 on one SM sub-partition (16 K registers) -> 128 per thread
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[DT_MAX_SLOTS], empty_bar[DT_MAX_SLOTS], tfull_bar[4], tempty_bar[4];
    __shared__ __align__(8) uint64_t bfull_bar[DT_MAX_BSTAGES], bempty_bar[DT_MAX_BSTAGES];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t key_s[32];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t NS = p.n_slots;
    uint8_t *toep = smem;
    uint8_t *ring = smem + DT_TOEP_BYTES;
    uint8_t *bring = ring + NS * DT_SLOT_BYTES;
    const uint32_t NB = p.n_bstages;

    if (tid < 32) key_s[tid] = p.key.w[tid];
    if (tid == 0) {
        for (uint32_t i = 0; i < NS; ++i) { mbar_init(&full_bar[i], DT_PROD_WARPS * 32u); mbar_init(&empty_bar[i], 1); }
        for (uint32_t i = 0; i < 4; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], DT_EPI_WARPS * 32u); }
        for (uint32_t i = 0; i < NB; ++i) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], DT_EPI_WARPS); }
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

    if (warp < DT_EPI_WARPS) {
        // ---------------- epilogue: phase = b - sum_l 2^(8l) P_l, decode, store. Warp = TMEM lane quadrant; every j block.
        const uint32_t qd = warp;
        const uint32_t j_w = qd * 32u + lane;        // + 128 jb
        const bool st_on = !(p.knockout & 2u);
        uint32_t pc = 0;                             // j blocks done: TMEM stage pc % 4, b stage pc % NB
        uint32_t bst = 0, bph = 0;
        for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
            const uint64_t ct0 = g * DT_CTS;
            const uint32_t n_here = (uint32_t) min((uint64_t) DT_CTS, p.n_ct - ct0);
#pragma unroll 1
            for (uint32_t jb = 0; jb < 8; ++jb, ++pc) {
                const uint32_t stage = pc & 3u;
                const uint32_t j = jb * 128u + j_w;
                const uint32_t taddr = tmem + ((qd * 32u) << 16) + stage * DT_N;
                float *sc = p.scores ? p.scores + ct0 * p.S + j : nullptr;
                uint32_t *ph = p.phase ? p.phase + ct0 * POLY_N + j : nullptr;
                mbar_wait(&tfull_bar[stage], (pc >> 2) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool sc_on = sc != nullptr && j < p.S && st_on;
                const uint32_t h0 = bst, p0 = bph;
                if (++bst == NB) { bst = 0; bph ^= 1u; }
                const uint32_t h1 = bst, p1 = bph;
                if (++bst == NB) { bst = 0; bph ^= 1u; }
                const uint32_t a0 = smem_u32(bring + h0 * DT_B_STAGE_BYTES) + 4u * j_w, a1 = smem_u32(bring + h1 * DT_B_STAGE_BYTES) + 4u * j_w;
                if (n_here == DT_CTS) dec_epilogue_pass<PHASE, true>(a0, a1, &bfull_bar[h0], &bfull_bar[h1], p0, p1, &bempty_bar[h0], &bempty_bar[h1], n_here, taddr, sc, p.S, sc_on, ph, &tempty_bar[stage]);
                else dec_epilogue_pass<PHASE, false>(a0, a1, &bfull_bar[h0], &bfull_bar[h1], p0, p1, &bempty_bar[h0], &bempty_bar[h1], n_here, taddr, sc, p.S, sc_on, ph, &tempty_bar[stage]);
            }
        }
    } else if (warp == DT_WARP_BLOAD) {
        // ---------------- b loader: ONE tensor copy (TMA, cp.async.bulk.tensor.2d) per half stage = the 128 b words of a j block of 16
        // ciphertexts, a 512-byte x 16-row box of the array {b polynomial of every ciphertext, row pitch = ciphertext stride}. As 32
        // per-ciphertext bulk copies per j block the issue loop alone (elect, broadcast, copy, branch per lane) took most of a pass and
        // the epilogue waited for b words 43 % of its time (ncu, profiles/r02_ncu_decrypt_tc.txt). Rows past the last ciphertext are
        // zero-filled by the copy and count towards the barrier's bytes like any other.
        if (lane == 0) {
            uint32_t bst = 0, bph = 0;
            const uint64_t map = reinterpret_cast<uint64_t>(&bmap);
            for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
                const uint32_t row0 = (uint32_t) (g * DT_CTS);
#pragma unroll 1
                for (uint32_t jb = 0; jb < 8; ++jb) {
#pragma unroll 1
                    for (uint32_t half = 0; half < 2; ++half) {
                        uint64_t *const bar = &bfull_bar[bst];
                        mbar_wait(&bempty_bar[bst], bph ^ 1u);
                        if (p.knockout & 4u) mbar_arrive(bar);
                        else {
                            mbar_arrive_expect_tx(bar, DT_B_STAGE_BYTES);
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                         ::"r"(smem_u32(bring + bst * DT_B_STAGE_BYTES)), "l"(map), "r"(jb * 128u), "r"(row0 + 16u * half), "r"(smem_u32(bar)) : "memory");
                        }
                        if (++bst == NB) { bst = 0; bph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == DT_WARP_MMA) {
        // ---------------- MMA issuer
        const uint32_t leader = elect_one();
        // descriptors (SWIZZLE_NONE, version 1): low word = address >> 4 | (LBO >> 4) << 16, high word = SBO >> 4 | 1 << 14. Every
        // operand address stays below 2^18, so stepping through the Toeplitz table / the ring is an add on the low word.
        const uint32_t a_lo0 = ((smem_u32(toep) >> 4) & 0x3FFFu) | ((DT_A_LBO >> 4) << 16), a_hi = (DT_A_SBO >> 4) | (1u << 14);
        const uint32_t b_lo0 = ((smem_u32(ring) >> 4) & 0x3FFFu) | ((DT_B_LBO >> 4) << 16), b_hi = (DT_B_SBO >> 4) | (1u << 14);
        auto mk = [](uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; };
        uint32_t slot0 = 0, ph0 = 0, pc = 0;       // pc counts j blocks: 8 per group, TMEM stage pc % 4
        if (!(p.knockout & 64u)) {
            // Weight-stationary schedule (the production schedule; knock-out 64 selects the one-j-block-at-a-time loop below): the j blocks are taken in PAIRS. For every K step the byte-plane chunk B (4 KB) is fetched
            // into the tensor core's collector buffer once and multiplied by the Toeplitz tiles of both j blocks (tcgen05.mma.ws,
            // fill -> lastuse), so an MMA reads 6 KB of operands from shared memory instead of 8 KB -- the pipe the producers' operand
            // stores and the epilogue share with the MMAs. Two accumulators (TMEM stages pc % 4 and pc % 4 + 1) are written per pair.
            for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
#pragma unroll 1
                for (uint32_t jp = 0; jp < 4; ++jp, pc += 2) {
                    const uint32_t st0 = pc & 3u;               // even: st0 + 1 <= 3
                    mbar_wait(&tempty_bar[st0], ((pc >> 2) & 1u) ^ 1u);
                    mbar_wait(&tempty_bar[st0 + 1u], ((pc >> 2) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t slot = slot0, ph = ph0;
                    uint32_t a_lo = a_lo0 + 128u * (2u * jp);
                    const uint32_t d0 = tmem + st0 * DT_N, d1 = d0 + DT_N;
#pragma unroll 1
                    for (uint32_t s = 0; s < DT_GROUP_SLOTS; ++s, a_lo += 128u) {
                        if (jp == 0) {
                            mbar_wait(&full_bar[slot], ph);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint32_t b_lo = b_lo0 + slot * (DT_SLOT_BYTES >> 4);
                        if (!(p.knockout & 1u))
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) {
                            const uint64_t db = mk(b_lo + kk * ((2u * DT_B_LBO) >> 4), b_hi);
                            tc_mma_ws_p<0, 0>(d0, mk(a_lo + 32u * kk, a_hi), db, dec_idesc(DT_N), (s | kk) != 0u, leader);
                            tc_mma_ws_p<0, 2>(d1, mk(a_lo + 128u + 32u * kk, a_hi), db, dec_idesc(DT_N), (s | kk) != 0u, leader);
                        }
                        if (jp == 3 && leader) tc_commit(&empty_bar[slot]);   // the group is done with this slot
                        if (++slot == NS) { slot = 0; ph ^= 1u; }
                    }
                    if (leader) { tc_commit(&tfull_bar[st0]); tc_commit(&tfull_bar[st0 + 1u]); }
                    __syncwarp();
                }
                slot0 += DT_GROUP_SLOTS;
                if (slot0 >= NS) { slot0 -= NS; ph0 ^= 1u; }
            }
        } else
        for (uint64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
#pragma unroll 1
            for (uint32_t jb = 0; jb < 8; ++jb, ++pc) {
                const uint32_t stage = pc & 3u;
                mbar_wait(&tempty_bar[stage], ((pc >> 2) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t slot = slot0, ph = ph0;
                uint32_t a_lo = a_lo0 + 128u * jb;            // A tile (jb, ks): + 32 ks  [16-byte units]
                const uint32_t d0 = tmem + stage * DT_N;
#pragma unroll 1
                for (uint32_t s = 0; s < DT_GROUP_SLOTS; ++s, a_lo += 128u) {
                    if (jb == 0) {
                        mbar_wait(&full_bar[slot], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t b_lo = b_lo0 + slot * (DT_SLOT_BYTES >> 4);
                    if (!(p.knockout & 1u))
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk)
                        tc_mma_p(d0, mk(a_lo + 32u * kk, a_hi), mk(b_lo + kk * ((2u * DT_B_LBO) >> 4), b_hi), dec_idesc(DT_N), (s | kk) != 0u, leader);
                    if (jb == 7 && leader) tc_commit(&empty_bar[slot]);   // the group is done with this slot
                    if (++slot == NS) { slot = 0; ph ^= 1u; }
                }
                if (leader) tc_commit(&tfull_bar[stage]);
                __syncwarp();
            }
            slot0 += DT_GROUP_SLOTS;
            if (slot0 >= NS) { slot0 -= NS; ph0 ^= 1u; }
        }
    } else {
        // ---------------- producers: unit i = (group, slot s). Warp w stages ciphertexts 4 w .. 4 w + 3 of the group: load q
        // of a unit reads the 512 contiguous bytes (8 k' blocks) of ciphertext 4 w + q; lane = (block ub = lane / 4, piece = lane % 4).
        const uint32_t pw = warp - DT_WARP_PROD;
        const uint32_t ub = lane >> 2, piece = lane & 3u;
        const uint64_t my_groups = p.n_groups > blockIdx.x ? (p.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint64_t total = my_groups * DT_GROUP_SLOTS;
        auto load_unit = [&](uint64_t i, uint4 (&w)[4]) {
            if (i >= total || (p.knockout & 8u)) return;
            const uint64_t g = blockIdx.x + (i >> 3) * gridDim.x;
            const uint32_t kb = (uint32_t) (i & 7u) * 8u + ub;     // k' block; its coefficients are 16 (63 - kb) .. + 15
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint64_t ct = g * DT_CTS + pw * 4u + q;
                if (ct >= p.n_ct) ct = p.n_ct - 1;     // tail group: a valid address; the epilogue ignores these columns
                w[q] = ldg128(p.in.words + ct * STRIDE + 64u * (63u - kb) + 16u * piece);
            }
        };
        uint32_t slot = 0, ph = 0;
        const bool odd = piece & 1u, hi = piece & 2u;
        // raw words of a unit -> operand rows, in place (registers only)
        auto transform_unit = [&](uint4 (&w)[4]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // byte planes of this lane's 4 coefficients (reversed), then a 4 x 4 transpose over the 4 lanes of the block:
                // lane `piece` ends with plane `piece` of all 16 coefficients
                uint32_t v0, v1, v2, v3;
                dec_split_rev(w[q], v0, v1, v2, v3);
                const uint32_t r0 = __shfl_xor_sync(0xFFFFFFFFu, odd ? v0 : v1, 1), r1 = __shfl_xor_sync(0xFFFFFFFFu, odd ? v2 : v3, 1);
                const uint32_t a0 = odd ? r0 : v0, a1 = odd ? v1 : r0, a2 = odd ? r1 : v2, a3 = odd ? v3 : r1;
                const uint32_t s0 = __shfl_xor_sync(0xFFFFFFFFu, hi ? a0 : a2, 2), s1 = __shfl_xor_sync(0xFFFFFFFFu, hi ? a1 : a3, 2);
                // t[i] = plane `piece` of the coefficients held by lane i of the block; row bytes run from the highest coefficient down
                const uint32_t t0 = hi ? s0 : a0, t1 = hi ? s1 : a1, t2 = hi ? a2 : s0, t3 = hi ? a3 : s1;
                w[q] = make_uint4(t3, t2, t1, t0);
            }
        };
        auto store_unit = [&](const uint4 (&row)[4]) {
            mbar_wait(&empty_bar[slot], ph ^ 1u);
            uint8_t *dst = ring + slot * DT_SLOT_BYTES + ub * DT_B_LBO + (pw * 16u + piece) * 16u;   // n = 4 ct + plane
            if (!(p.knockout & 16u)) {
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4 *>(dst + q * 64u) = row[q];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_bar[slot]);
            if (++slot == NS) { slot = 0; ph ^= 1u; }
        };
        // Six units (4 x 16-byte loads each) in flight per thread, register sets bound statically; a unit goes load -> (four units
        // later) split + transpose in registers -> (two units later) store, as soon as the MMAs of the previous group have let go of
        // its slot. A group's slots come free only during its last pass over them, one every ~512 cycles, and what a producer warp
        // does between two stores has to fit in there: with the split + transpose (~160 instructions of one warp) done right before
        // the store, the first pass of every group waited for its operands ~19 % of the MMA warp's time (ncu,
        // profiles/r02_ncu_decrypt_tc.txt); done two units ahead, a freed slot is filled by a wait, four stores and an arrive.
        uint4 w0[4], w1[4], w2[4], w3[4], w4[4], w5[4];
        load_unit(0, w0);
        load_unit(1, w1);
        load_unit(2, w2);
        load_unit(3, w3);
        load_unit(4, w4);
        load_unit(5, w5);
        transform_unit(w0);
        transform_unit(w1);
        for (uint64_t i = 0; i < total; i += 6) {
            store_unit(w0);
            load_unit(i + 6, w0);
            transform_unit(w2);
            if (i + 1 < total) { store_unit(w1); load_unit(i + 7, w1); }
            transform_unit(w3);
            if (i + 2 < total) { store_unit(w2); load_unit(i + 8, w2); }
            transform_unit(w4);
            if (i + 3 < total) { store_unit(w3); load_unit(i + 9, w3); }
            transform_unit(w5);
            if (i + 4 < total) { store_unit(w4); load_unit(i + 10, w4); }
            transform_unit(w0);
            if (i + 5 < total) { store_unit(w5); load_unit(i + 11, w5); }
            transform_unit(w1);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == DT_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
