// K4p: the decrypt GEMM of decrypt_tc.cuh on CTA PAIRS -- tcgen05.mma.cta_group::2, M = 256.
//
// Same arithmetic, operand layouts and roles as K4t (read its header first). What changes is who holds what:
//   * a cluster of two CTAs (two SMs of one TPC) works on one group of 32 ciphertexts at a time. One MMA covers TWO j blocks
//     (M = 256: rows 0-127 = j block 2 jp from the Toeplitz table of CTA 0, rows 128-255 = j block 2 jp + 1 from the table of
//     CTA 1, which is built 128 coefficients further on so that both CTAs use the same descriptor) against all 32 ciphertexts
//     (N = 128), and each CTA holds HALF of the B operand: the byte planes of 16 ciphertexts, 8.5 KB per 128-coefficient slot instead
//     of 16.5 KB. A group is 8 slots = 68 KB per CTA, so TWO groups are resident (16 slots) where K4t has room for one group and two
//     slots: K4t's producers can refill only while a group makes its last pass over its slots (a quarter of its time, one slot every
//     ~512 cycles), here the whole next group is staged while the current one is multiplied.
//   * the accumulator of a pass is 128 TMEM lanes x 128 columns in EACH CTA (its own j block, all 32 ciphertexts); each CTA's
//     epilogue warps read their own TMEM, fetch the b words of their own j block (tensor copies) and store their own scores. A group is
//     4 passes; the 4 TMEM stages let the MMAs run three passes ahead of the epilogues.
//   * only CTA 0 issues MMAs. Its barriers collect both CTAs: full[slot] counts the 16 producer warps of the pair, t_empty[stage] the 8
//     epilogue warps (CTA 1's arrive through the cluster address: mapa + mbarrier.arrive.shared::cluster); tcgen05.commit
//     .multicast::cluster signals empty[slot] and t_full[stage] in both CTAs at once.
#pragma once

#define DP_B_LBO (16u * 64u + 64u)        // bytes between 16-k' column blocks of a slot: 64 n rows x 16 bytes + the conflict pad (1088 = 64 mod 128)
#define DP_SLOT_BYTES (8u * DP_B_LBO)     // one ring slot of one CTA: 128 k' x 64 n bytes = 4 K steps (8704 with the pads)
#define DP_MAX_SLOTS 20u

__host__ __device__ constexpr uint32_t dec_pair_smem_bytes(uint32_t n_slots, uint32_t n_bstages) {
    return DT_TOEP_BYTES + n_slots * DP_SLOT_BYTES + n_bstages * DT_B_STAGE_BYTES;
}
// instruction descriptor: D = s32, A = s8 MN-major, B = u8 K-major, M = 256 (two CTAs), N = n
__host__ __device__ constexpr uint32_t dec_pair_idesc(uint32_t n) {
    return (2u << 4) | (1u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrival on a barrier of the other CTA (or of this one) through its cluster address. Default semantics -- release at CTA scope -- as in
// CUTLASS's ClusterBarrier::arrive(cta_id): shared memory is not cached, so what the arriving thread wrote to its own shared memory is
// in place when the arrival reaches the other SM. The cluster-scope forms are much more than is needed here and very expensive:
// mbarrier.arrive.release.cluster waits for EVERY earlier memory operation of its thread -- the epilogue's score stores on their way
// to HBM (3800 instead of 1700 cycles per pass), the producers' seven units of loads in flight (memory side of the kernel 1.15
// instead of 0.5 ms) -- and mbarrier.try_wait.acquire.cluster costs the MMA thread ~1000 cycles per wait even when satisfied.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t) 3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair_p(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One pass of one epilogue warp (dec_epilogue_pass of K4t with the pair's hand-back: ONE arrival per warp on CTA 0's t_empty barrier,
// through its cluster address, once the warp's last TMEM load has completed)
template <bool PHASE, bool FULL>
__device__ __forceinline__ void dec_pair_epilogue_pass(uint32_t b_addr0, uint32_t b_addr1, uint64_t *bfull0, uint64_t *bfull1, uint32_t bpar0, uint32_t bpar1,
                                                       uint64_t *bempty0, uint64_t *bempty1, uint32_t n_here, uint32_t taddr, float *sc, uint32_t S,
                                                       bool sc_on, uint32_t *ph, uint32_t tempty_cluster) {
    uint8_t *scp = reinterpret_cast<uint8_t *>(sc);
    const uint32_t s4 = 4u * S;
    uint32_t vv[2][4][8];
#pragma unroll
    for (uint32_t m = 0; m < 4; ++m) tc_ld8(taddr + m * 8u, vv[0][m]);
#pragma unroll
    for (uint32_t chunk = 0; chunk < 4; ++chunk) {
        uint32_t (&v)[4][8] = vv[chunk & 1u];
        uint32_t bw[8];
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
        if (chunk & 1u) {
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive(chunk == 1 ? bempty0 : bempty1);
        }
        if (chunk == 3) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive_cluster(tempty_cluster);
        }
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const uint32_t cc = chunk * 8u + c;
            const uint32_t *q = &v[c >> 1][4u * (c & 1u)];
            const uint32_t phs = bw[c] - (q[0] + (q[1] << 8) + (q[2] << 16) + (q[3] << 24));
            const bool here = FULL || cc < n_here;
            if (PHASE && here) stg32_stream(ph + cc * POLY_N, phs);
            const uint32_t fbits = __float_as_uint(__int2float_rn((int32_t) phs) * 2.3283064365386963e-10f);
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.u32 [%0], %1;\n\t}\n" ::"l"(scp), "r"(fbits), "r"((uint32_t) (sc_on && here)) : "memory");
            scp += s4;
        }
    }
}

template <uint32_t STRIDE, bool PHASE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DT_THREADS, 1) decrypt_pair_kernel(const DecTcParams p, const __grid_constant__ CUtensorMap bmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[DP_MAX_SLOTS], empty_bar[DP_MAX_SLOTS], tfull_bar[4], tempty_bar[4];
    __shared__ __align__(8) uint64_t bfull_bar[DT_MAX_BSTAGES], bempty_bar[DT_MAX_BSTAGES];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t key_s[32];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t rank = cluster_ctarank();
    const uint64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t NS = p.n_slots, NB = p.n_bstages;
    uint8_t *toep = smem;
    uint8_t *ring = smem + DT_TOEP_BYTES;
    uint8_t *bring = ring + NS * DP_SLOT_BYTES;

    if (tid < 32) key_s[tid] = p.key.w[tid];
    if (tid == 0) {
        for (uint32_t i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 2u * DT_PROD_WARPS); mbar_init(&empty_bar[i], 1); }
        for (uint32_t i = 0; i < 4; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 2u * DT_EPI_WARPS); }
        for (uint32_t i = 0; i < NB; ++i) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], DT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == DT_WARP_MMA) {      // the same warp of both CTAs, the same destination word
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    __syncthreads();
    // Toeplitz table of this CTA's j blocks: core matrix i, row kk (k' offset), byte mm (j offset) = t(8 i + kk + mm - 1023 + 128 rank)
    for (uint32_t idx = tid * 4u; idx < DT_TOEP_CORES * 128u; idx += DT_THREADS * 4u) {
        const int32_t x0 = (int32_t) (8u * (idx >> 7) + ((idx >> 4) & 7u) + (idx & 15u) + 128u * rank) - 1023;
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int32_t x = x0 + b;
            if (x > 1023) continue;      // past the last core matrix a descriptor of this CTA reaches (odd j blocks end 128 earlier)
            const uint32_t xi = (uint32_t) (x >= 0 ? x : x + 1024);
            const uint32_t bit = (key_s[xi >> 5] >> (xi & 31u)) & 1u;
            const uint32_t v = x >= 0 ? bit : (0u - bit) & 0xFFu;
            word |= v << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(toep + idx) = word;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();       // both CTAs: barriers initialised, tables built, TMEM allocated -- before any remote arrival or MMA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp < DT_EPI_WARPS) {
        // ---------------- epilogue of this CTA's j block of every pass: j block 2 jp + rank
        const uint32_t qd = warp;
        const uint32_t j_w = qd * 32u + lane;
        const bool st_on = !(p.knockout & 2u);
        const uint32_t tempty0 = mapa_u32(smem_u32(&tempty_bar[0]), 0u);      // CTA 0's t_empty barriers
        uint32_t pc = 0, bst = 0, bph = 0;
        for (uint64_t g = pair; g < p.n_groups; g += n_pairs) {
            const uint64_t ct0 = g * DT_CTS;
            const uint32_t n_here = (uint32_t) min((uint64_t) DT_CTS, p.n_ct - ct0);
#pragma unroll 1
            for (uint32_t jp = 0; jp < 4; ++jp, ++pc) {
                const uint32_t stage = pc & 3u;
                const uint32_t j = (2u * jp + rank) * 128u + j_w;
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
                if (n_here == DT_CTS) dec_pair_epilogue_pass<PHASE, true>(a0, a1, &bfull_bar[h0], &bfull_bar[h1], p0, p1, &bempty_bar[h0], &bempty_bar[h1], n_here, taddr, sc, p.S, sc_on, ph, tempty0 + 8u * stage);
                else dec_pair_epilogue_pass<PHASE, false>(a0, a1, &bfull_bar[h0], &bfull_bar[h1], p0, p1, &bempty_bar[h0], &bempty_bar[h1], n_here, taddr, sc, p.S, sc_on, ph, tempty0 + 8u * stage);
            }
        }
    } else if (warp == DT_WARP_BLOAD) {
        // ---------------- b loader: the 128 b words of this CTA's j block, 16 ciphertexts per tensor copy (as in K4t)
        if (lane == 0) {
            uint32_t bst = 0, bph = 0;
            const uint64_t map = reinterpret_cast<uint64_t>(&bmap);
            for (uint64_t g = pair; g < p.n_groups; g += n_pairs) {
                const uint32_t row0 = (uint32_t) (g * DT_CTS);
#pragma unroll 1
                for (uint32_t jp = 0; jp < 4; ++jp) {
#pragma unroll 1
                    for (uint32_t half = 0; half < 2; ++half) {
                        uint64_t *const bar = &bfull_bar[bst];
                        mbar_wait(&bempty_bar[bst], bph ^ 1u);
                        if (p.knockout & 4u) mbar_arrive(bar);
                        else {
                            mbar_arrive_expect_tx(bar, DT_B_STAGE_BYTES);
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                         ::"r"(smem_u32(bring + bst * DT_B_STAGE_BYTES)), "l"(map), "r"((2u * jp + rank) * 128u), "r"(row0 + 16u * half), "r"(smem_u32(bar)) : "memory");
                        }
                        if (++bst == NB) { bst = 0; bph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == DT_WARP_MMA) {
        // ---------------- MMA issuer: CTA 0 only, for the pair
        if (rank == 0) {
            const uint32_t leader = elect_one();
            const uint32_t a_lo0 = ((smem_u32(toep) >> 4) & 0x3FFFu) | ((DT_A_LBO >> 4) << 16), a_hi = (DT_A_SBO >> 4) | (1u << 14);
            const uint32_t b_lo0 = ((smem_u32(ring) >> 4) & 0x3FFFu) | ((DP_B_LBO >> 4) << 16), b_hi = (DT_B_SBO >> 4) | (1u << 14);
            auto mk = [](uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; };
            uint32_t slot0 = 0, ph0 = 0, pc = 0;       // pc counts passes: 4 per group, TMEM stage pc % 4
            for (uint64_t g = pair; g < p.n_groups; g += n_pairs) {
#pragma unroll 1
                for (uint32_t jp = 0; jp < 4; ++jp, ++pc) {
                    const uint32_t stage = pc & 3u;
                    mbar_wait(&tempty_bar[stage], ((pc >> 2) & 1u) ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t slot = slot0, ph = ph0;
                    uint32_t a_lo = a_lo0 + 256u * jp;            // A tile (j block 2 jp, K step ks): + 32 ks  [16-byte units]
                    const uint32_t d0 = tmem + stage * DT_N;
#pragma unroll 1
                    for (uint32_t s = 0; s < DT_GROUP_SLOTS; ++s, a_lo += 128u) {
                        if (jp == 0) {
                            mbar_wait(&full_bar[slot], ph);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint32_t b_lo = b_lo0 + slot * (DP_SLOT_BYTES >> 4);
                        if (!(p.knockout & 1u))
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk)
                            tc_mma_pair_p(d0, mk(a_lo + 32u * kk, a_hi), mk(b_lo + kk * ((2u * DP_B_LBO) >> 4), b_hi), dec_pair_idesc(DT_N), (s | kk) != 0u, leader);
                        if (jp == 3 && leader) tc_commit_pair(&empty_bar[slot]);   // the group is done with this slot, in both CTAs
                        if (++slot == NS) { slot = 0; ph ^= 1u; }
                    }
                    if (leader) tc_commit_pair(&tfull_bar[stage]);
                    __syncwarp();
                }
                slot0 += DT_GROUP_SLOTS;
                if (slot0 >= NS) { slot0 -= NS; ph0 ^= 1u; }
            }
        }
    } else {
        // ---------------- producers: unit i = (group, slot s). Warp w stages ciphertexts 16 rank + 2 w, + 1 of the group (K4t: 4 w .. + 3);
        // lane = (block ub = lane / 4, piece = lane % 4)
        const uint32_t pw = warp - DT_WARP_PROD;
        const uint32_t ub = lane >> 2, piece = lane & 3u;
        const uint64_t my_groups = p.n_groups > pair ? (p.n_groups - pair + n_pairs - 1) / n_pairs : 0;
        const uint64_t total = my_groups * DT_GROUP_SLOTS;
        const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0u);        // CTA 0's full barriers
        auto load_unit = [&](uint64_t i, uint4 (&w)[2]) {
            if (i >= total || (p.knockout & 8u)) return;
            const uint64_t g = pair + (i >> 3) * n_pairs;
            const uint32_t kb = (uint32_t) (i & 7u) * 8u + ub;     // k' block; its coefficients are 16 (63 - kb) .. + 15
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint64_t ct = g * DT_CTS + 16u * rank + pw * 2u + q;
                if (ct >= p.n_ct) ct = p.n_ct - 1;     // tail group: a valid address; the epilogue ignores these columns
                w[q] = ldg128(p.in.words + ct * STRIDE + 64u * (63u - kb) + 16u * piece);
            }
        };
        uint32_t slot = 0, ph = 0;
        const bool odd = piece & 1u, hi = piece & 2u;
        auto transform_unit = [&](uint4 (&w)[2]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t v0, v1, v2, v3;
                dec_split_rev(w[q], v0, v1, v2, v3);
                const uint32_t r0 = __shfl_xor_sync(0xFFFFFFFFu, odd ? v0 : v1, 1), r1 = __shfl_xor_sync(0xFFFFFFFFu, odd ? v2 : v3, 1);
                const uint32_t a0 = odd ? r0 : v0, a1 = odd ? v1 : r0, a2 = odd ? r1 : v2, a3 = odd ? v3 : r1;
                const uint32_t s0 = __shfl_xor_sync(0xFFFFFFFFu, hi ? a0 : a2, 2), s1 = __shfl_xor_sync(0xFFFFFFFFu, hi ? a1 : a3, 2);
                const uint32_t t0 = hi ? s0 : a0, t1 = hi ? s1 : a1, t2 = hi ? a2 : s0, t3 = hi ? a3 : s1;
                w[q] = make_uint4(t3, t2, t1, t0);
            }
        };
        auto store_unit = [&](const uint4 (&row)[2]) {
            mbar_wait(&empty_bar[slot], ph ^ 1u);
            uint8_t *dst = ring + slot * DP_SLOT_BYTES + ub * DP_B_LBO + (pw * 8u + piece) * 16u;   // n = 4 (2 w + q) + plane
            if (!(p.knockout & 16u)) {
#pragma unroll
                for (int q = 0; q < 2; ++q) *reinterpret_cast<uint4 *>(dst + q * 64u) = row[q];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(full0 + 8u * slot);       // one arrival per producer warp, on CTA 0's barrier
            if (++slot == NS) { slot = 0; ph ^= 1u; }
        };
        // eight units (2 x 16-byte loads each) in flight per thread = a whole group; a unit goes load -> (five units later) split +
        // transpose in registers -> (three units later) wait for its slot, two stores, arrive
        uint4 w0[2], w1[2], w2[2], w3[2], w4[2], w5[2], w6[2], w7[2];
        load_unit(0, w0); load_unit(1, w1); load_unit(2, w2); load_unit(3, w3);
        load_unit(4, w4); load_unit(5, w5); load_unit(6, w6); load_unit(7, w7);
        transform_unit(w0); transform_unit(w1); transform_unit(w2);
        for (uint64_t i = 0; i < total; i += 8) {      // total is a multiple of 8
            store_unit(w0); load_unit(i + 8, w0); transform_unit(w3);
            store_unit(w1); load_unit(i + 9, w1); transform_unit(w4);
            store_unit(w2); load_unit(i + 10, w2); transform_unit(w5);
            store_unit(w3); load_unit(i + 11, w3); transform_unit(w6);
            store_unit(w4); load_unit(i + 12, w4); transform_unit(w7);
            store_unit(w5); load_unit(i + 13, w5); transform_unit(w0);
            store_unit(w6); load_unit(i + 14, w6); transform_unit(w1);
            store_unit(w7); load_unit(i + 15, w7); transform_unit(w2);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();       // neither CTA leaves (its shared memory, its barriers) while the other may still reach into it
    if (warp == DT_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
