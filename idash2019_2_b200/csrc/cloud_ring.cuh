// K2r: persistent, warp-specialised variant of the tensor-core cloud kernel for NUM_REGIONS == 1
// (the iDASH-scale configurations). Same arithmetic as cloud_tc.cuh; what changes is the schedule:
//
//   * grid = 16 slices x C chunks (C = SMs / 16): CTA (slice, chunk) walks the tiles of its chunk in order
//     for ONE 128-word slice of the ciphertext axis. The 16 CTAs of a chunk advance together, so whole
//     8 KB ciphertexts are read / written at about the same time and a tile's coefficient image is
//     fetched from HBM once and served to the other 15 CTAs by L2.
//   * the limb planes of the input blocks (32 features x 128 words x 4 planes = 18 KB) live in a
//     shared-memory RING: consecutive tiles share most of their band, so every (block, slice) is loaded
//     from global memory, split into byte planes and stored ONCE per CTA -- the north-star's "stage each
//     overlapping tag window once and reuse it across consecutive target SNPs".
//   * roles: warps 0-7 epilogue (TMEM lane quadrant = warp id % 4, rows 32 (warp id / 4) .. +31 of the tile),
//     warp 8 MMA issuer (+ TMEM owner), warp 9 coefficient loader (cp.async.bulk, one 4 KB chunk per
//     32-feature K step into its own ring), warps 10-13 block producers. mbarrier pipelines: a_full/a_empty
//     per input-block slot, b_full/b_empty per coefficient-chunk slot, t_full/t_empty x2 TMEM stages
//     (2 x 4 accumulators x 64 columns = all 512 TMEM columns), so the MMAs of tile t+1 overlap the
//     epilogue of tile t, the coefficient loader runs several tiles ahead and the producers a ring ahead.
#pragma once

#define RG_THREADS 448
#define RG_EPI_WARPS 8
#define RG_WARP_MMA 8
#define RG_WARP_BLOAD 9
#define RG_WARP_PROD 10
#define RG_BLOCK_BYTES (16u * TC_A_LBO)          // 4 planes x 4 feature groups x 1152 B = 18432
#define RG_PLANE_BYTES (4u * TC_A_LBO)
#define RG_MAX_SLOTS 9
#define RG_MAX_BSLOTS 16
#define RG_SMEM_MAX (226u * 1024u)               // dynamic shared memory budget of the one resident CTA

struct RingParams {
    const idash_b200_tile *tiles;
    const uint32_t *tile_rows;
    const int32_t *tile_bias;
    const uint8_t *tile_coef;
    const uint32_t *feat_used;     // bit f: feature f is used by some row
    uint32_t n_feat_words;
    uint32_t n_tiles;
    uint32_t n_chunks;             // gridDim.x = 16 * n_chunks
    uint32_t n_slots;              // input-block ring slots (>= widest tile in blocks, + prefetch)
    uint32_t n_bslots;             // coefficient-chunk ring slots
    CtView in, out;
    const uint32_t *slot_of_ct;
    uint32_t n_ct_slots;
    const uint32_t *slot_of_row;
    uint32_t S;
    int *status;
    uint32_t trace_cta;            // 0 = off, else 1 + index of the CTA whose timeline is recorded
    uint32_t knockout;             // profiling aid (IDASH_B200_KNOCKOUT, results are wrong when non-zero):
                                   // 1 no MMAs, 2 no output stores, 4 no epilogue TMEM loads, 8 no input loads, 16 no coefficient copies
};

__host__ __device__ constexpr uint32_t ring_smem_bytes(uint32_t n_slots, uint32_t n_bslots) { return n_slots * RG_BLOCK_BYTES + n_bslots * TC_B_CHUNK; }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tracing aid (IDASH_B200_TRACE=<cta>): per-tile SM-clock timestamps of one CTA, see tools/trace_ring.py
#define RG_TRACE_TILES 96
#define RG_TRACE_EVENTS 8
__device__ unsigned long long g_ring_trace[RG_TRACE_TILES * RG_TRACE_EVENTS];
#define RG_TRACE(ev, it_)                                                                                     \
    do {                                                                                                      \
        if (p.trace_cta == blockIdx.x + 1u && (it_) >= 64u && (it_) < 64u + RG_TRACE_TILES)                     \
            g_ring_trace[((it_) - 64u) * RG_TRACE_EVENTS + (ev)] = clock64();                                    \
    } while (0)

struct RingTile { uint32_t a, nb; uint64_t b_off; };   // first block, blocks, coefficient image

__device__ __forceinline__ RingTile ring_tile(const RingParams &p, uint32_t t) {
    const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(p.tiles + t));
    RingTile r;
    r.a = t0.x >> 5; r.nb = t0.y >> 5;
    r.b_off = (uint64_t) t0.z | ((uint64_t) t0.w << 32);
    return r;
}

// Tile headers come from global memory (~1 us away): every role reads them RG_AHEAD tiles ahead of use through a
// small register queue, otherwise each tile iteration would expose one dependent L2 / HBM round trip.
#define RG_AHEAD 4
struct RingTileQueue {
    RingTile q[RG_AHEAD];
    uint32_t t_end;
    __device__ __forceinline__ void init(const RingParams &p, uint32_t t_begin, uint32_t t_end_) {
        t_end = t_end_;
#pragma unroll
        for (uint32_t i = 0; i < RG_AHEAD; ++i) q[i] = ring_tile(p, min(t_begin + i, t_end_ - 1));
    }
    // header of tile t (front) is consumed; fetch tile t + RG_AHEAD
    __device__ __forceinline__ void pop(const RingParams &p, uint32_t t) {
#pragma unroll
        for (uint32_t i = 0; i + 1 < RG_AHEAD; ++i) q[i] = q[i + 1];
        q[RG_AHEAD - 1] = ring_tile(p, min(t + RG_AHEAD, t_end - 1));
    }
};

__device__ __forceinline__ void ring_ld_chunk(uint32_t taddr, uint32_t (&v0)[8], uint32_t (&v1)[8], uint32_t (&v2)[8], uint32_t (&v3)[8]) {
    tc_ld8(taddr, v0);
    tc_ld8(taddr + TC_TN, v1);
    tc_ld8(taddr + 2 * TC_TN, v2);
    tc_ld8(taddr + 3 * TC_TN, v3);
}

// Epilogue of one tile for one warp: 32 words (TMEM lanes) x 32 rows (columns col_base .. +31 of each of the 4
// accumulators), 8 rows per tcgen05.ld group, the loads of group g+1 in flight while group g is recombined and
// stored. FAST: rows are consecutive output slots with a compile-time stride -> store address = base + immediate.
// ptr_own / bias_own: lane l holds the output address (0 = no such row) / Constant * 2^18 of row col_base + l.
template <bool FAST, uint32_t STRIDE, bool BIAS>
__device__ __forceinline__ void ring_epilogue(uint32_t taddr, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_own, uint32_t bias_flag,
                                              uint32_t lane_off, uint32_t knockout) {
    if (knockout & 4u) return;
    uint32_t v[2][4][8];
    ring_ld_chunk(taddr, v[0][0], v[0][1], v[0][2], v[0][3]);
#pragma unroll
    for (uint32_t g = 0; g < 4; ++g) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (g + 1 < 4) ring_ld_chunk(taddr + (g + 1) * 8u, v[(g + 1) & 1][0], v[(g + 1) & 1][1], v[(g + 1) & 1][2], v[(g + 1) & 1][3]);
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const uint32_t n = g * 8u + c;
            uint32_t x = ((v[g & 1][3][c] * 256u + v[g & 1][2][c]) * 256u + v[g & 1][1][c]) * 256u + v[g & 1][0][c];
            if (BIAS) x += __shfl_sync(0xFFFFFFFFu, bias_own, n) * bias_flag;
            if (knockout & 2u) { if (x == 0x9E3779B9u && bias_flag == 77u) stg32_stream(base_lane, x); continue; }
            if (FAST) {
                stg32_stream(base_lane + (uint64_t) n * STRIDE, x);
            } else {
                const uint64_t ptr = __shfl_sync(0xFFFFFFFFu, ptr_own, n);
                if (ptr) stg32_stream(reinterpret_cast<uint8_t *>(ptr) + lane_off, x);
            }
        }
    }
}

__global__ void __launch_bounds__(RG_THREADS, 1) cloud_ring_kernel(const RingParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[RG_MAX_SLOTS], a_empty[RG_MAX_SLOTS], b_full[RG_MAX_BSLOTS], b_empty[RG_MAX_BSLOTS], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t slice = blockIdx.x & 15u, chunk = blockIdx.x >> 4;
    const uint32_t t_begin = (uint32_t) ((uint64_t) p.n_tiles * chunk / p.n_chunks);
    const uint32_t t_end = (uint32_t) ((uint64_t) p.n_tiles * (chunk + 1) / p.n_chunks);
    if (t_begin >= t_end) return;
    uint8_t *sA = smem;
    uint8_t *sB = smem + p.n_slots * RG_BLOCK_BYTES;
    const uint32_t w_slice = slice * 128u;
    const bool is_b = (w_slice & POLY_N) != 0;
    const uint32_t i_slice = w_slice & (POLY_N - 1);

    if (tid == 0) {
        for (uint32_t s = 0; s < p.n_slots; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (uint32_t s = 0; s < p.n_bslots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (uint32_t s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], RG_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == RG_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp < RG_EPI_WARPS) {
        // ================= epilogue =================
        const uint32_t quad = warp & 3u, col_base = (warp >> 2) * 32u;
        const uint32_t word_in_slice = quad * 32u + lane;
        const uint32_t lane_off = 4u * word_in_slice;
        const uint32_t bias_flag = (is_b && i_slice + word_in_slice < p.S) ? 1u : 0u;
        const bool records = p.out.records != 0;
        uint32_t it = 0;
        // row information is fetched RG_AHEAD tiles ahead: lane l owns row col_base + l
        uint32_t row_q[RG_AHEAD], flags_q[RG_AHEAD];
        int32_t bias_q[RG_AHEAD];
#pragma unroll
        for (uint32_t i = 0; i < RG_AHEAD; ++i) {
            const uint32_t tt = min(t_begin + i, t_end - 1);
            row_q[i] = __ldg(p.tile_rows + (uint64_t) tt * TC_TN + col_base + lane);
            bias_q[i] = __ldg(p.tile_bias + (uint64_t) tt * TC_TN + col_base + lane);
            flags_q[i] = __ldg(&p.tiles[tt].flags);
        }
        for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
            const uint32_t row = row_q[0], flags = flags_q[0];
            const uint32_t bias_own = (uint32_t) bias_q[0] * (uint32_t) IDASH_B200_ONE_IN_T32;
            {
#pragma unroll
                for (uint32_t i = 0; i + 1 < RG_AHEAD; ++i) { row_q[i] = row_q[i + 1]; bias_q[i] = bias_q[i + 1]; flags_q[i] = flags_q[i + 1]; }
                const uint32_t tt = min(t + RG_AHEAD, t_end - 1);
                row_q[RG_AHEAD - 1] = __ldg(p.tile_rows + (uint64_t) tt * TC_TN + col_base + lane);
                bias_q[RG_AHEAD - 1] = __ldg(p.tile_bias + (uint64_t) tt * TC_TN + col_base + lane);
                flags_q[RG_AHEAD - 1] = __ldg(&p.tiles[tt].flags);
            }
            const bool fast = (flags & 1u) && p.slot_of_row == nullptr;
            uint64_t ptr_own = 0;
            uint8_t *base_lane = nullptr;
            if (fast) {
                const uint32_t row0 = __shfl_sync(0xFFFFFFFFu, row, 0);     // caller row of tile row col_base
                base_lane = p.out.words + (uint64_t) row0 * p.out.stride + 4u * w_slice + lane_off;
            } else if (row != IDASH_B200_NO_ROW) {
                ptr_own = (uint64_t) (p.out.words + (uint64_t) (p.slot_of_row ? __ldg(p.slot_of_row + row) : row) * p.out.stride + 4u * w_slice);
            }
            const uint32_t st = it & 1u;
            if (tid == 0) RG_TRACE(5, it);
            mbar_wait(&t_full[st], (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) RG_TRACE(6, it);
            const uint32_t taddr = tmem + ((quad * 32u) << 16) + st * 4u * TC_TN + col_base;
            if (fast) {
                if (records) {
                    if (is_b) ring_epilogue<true, IDASH_B200_RECORD_BYTES, true>(taddr, base_lane, 0, bias_own, bias_flag, lane_off, p.knockout);
                    else ring_epilogue<true, IDASH_B200_RECORD_BYTES, false>(taddr, base_lane, 0, 0, 0, lane_off, p.knockout);
                } else {
                    if (is_b) ring_epilogue<true, IDASH_B200_CT_BYTES, true>(taddr, base_lane, 0, bias_own, bias_flag, lane_off, p.knockout);
                    else ring_epilogue<true, IDASH_B200_CT_BYTES, false>(taddr, base_lane, 0, 0, 0, lane_off, p.knockout);
                }
            } else {
                ring_epilogue<false, 0, true>(taddr, nullptr, ptr_own, bias_own, bias_flag, lane_off, p.knockout);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (tid == 0) RG_TRACE(7, it);
            if (lane == 0) mbar_arrive(&t_empty[st]);      // one arrival per epilogue warp
        }
    } else if (warp == RG_WARP_MMA) {
        // ================= MMA issuer =================
        // Ring positions are tracked incrementally (no divisions); descriptors are a precomputed constant plus
        // a 16-byte-unit offset, so one k-step (7 MMAs) is a few dozen instructions for the issuing lane.
        const uint32_t n_slots = p.n_slots, n_bslots = p.n_bslots;
        const uint64_t da_base = tc_desc(smem_u32(sA), TC_A_LBO, TC_A_SBO);
        const uint64_t db_base = tc_desc(smem_u32(sB), TC_B_LBO, TC_B_SBO);
        uint32_t it = 0;
        uint32_t next_slot = 0, next_par = 0;       // slot / phase parity of the next input block to be staged
        uint32_t first_slot = 0;                    // slot of block T.a
        uint32_t bslot = 0, bpar = 0;               // coefficient-chunk ring position
        RingTileQueue tq;
        tq.init(p, t_begin, t_end);
        uint32_t staged_upto = tq.q[0].a, rel_upto = tq.q[0].a;
        for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
            const RingTile T = tq.q[0], Tn = tq.q[1];
            const bool has_next = t + 1 < t_end;
            tq.pop(p, t);
            const uint32_t bt = T.a + T.nb;
            const uint32_t st = it & 1u;
            if (lane == 0) RG_TRACE(0, it);
            mbar_wait(&t_empty[st], ((it >> 1) & 1u) ^ 1u);
            if (lane == 0) RG_TRACE(1, it);
            // input blocks this tile adds to the ring
            for (uint32_t kb = max(T.a, staged_upto); kb < bt; ++kb) {
                mbar_wait(&a_full[next_slot], next_par);
                if (++next_slot == n_slots) { next_slot = 0; next_par ^= 1u; }
            }
            staged_upto = max(staged_upto, bt);
            if (lane == 0) RG_TRACE(2, it);
            // coefficient chunks of this tile
            {
                uint32_t s = bslot, par = bpar;
                for (uint32_t ks = 0; ks < T.nb; ++ks) {
                    mbar_wait(&b_full[s], par);
                    if (++s == n_bslots) { s = 0; par ^= 1u; }
                }
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t rel_end = has_next ? min(Tn.a, bt) : bt;
            if (lane == 0) RG_TRACE(3, it);
            if (lane == 0) {
                const uint32_t d0 = tmem + st * 4u * TC_TN;
                uint32_t aslot = first_slot;
                for (uint32_t ks = 0; ks < T.nb; ++ks) {
                    const uint64_t da = da_base + (uint64_t) ((aslot * RG_BLOCK_BYTES) >> 4);
                    const uint64_t db = db_base + (uint64_t) ((bslot * TC_B_CHUNK) >> 4);
                    const uint32_t acc = ks ? 1u : 0u;
                    if (!(p.knockout & 1u)) {
                    tc_mma(d0 + 0 * TC_TN, da + 0 * (RG_PLANE_BYTES >> 4), db, tc_idesc(0), acc);                            // P0  = X0 c_lo
                    tc_mma(d0 + 1 * TC_TN, da + 0 * (RG_PLANE_BYTES >> 4), db + (TC_B_CHUNK >> 5), tc_idesc(1), acc);        // P1  = X0 c_hi
                    tc_mma(d0 + 1 * TC_TN, da + 1 * (RG_PLANE_BYTES >> 4), db, tc_idesc(0), 1u);                             // P1 += X1 c_lo
                    tc_mma(d0 + 2 * TC_TN, da + 1 * (RG_PLANE_BYTES >> 4), db + (TC_B_CHUNK >> 5), tc_idesc(1), acc);        // P2  = X1 c_hi
                    tc_mma(d0 + 2 * TC_TN, da + 2 * (RG_PLANE_BYTES >> 4), db, tc_idesc(0), 1u);                             // P2 += X2 c_lo
                    tc_mma(d0 + 3 * TC_TN, da + 2 * (RG_PLANE_BYTES >> 4), db + (TC_B_CHUNK >> 5), tc_idesc(1), acc);        // P3  = X2 c_hi
                    tc_mma(d0 + 3 * TC_TN, da + 3 * (RG_PLANE_BYTES >> 4), db, tc_idesc(0), 1u);                             // P3 += X3 c_lo
                    }
                    tc_commit(&b_empty[bslot]);
                    if (++aslot == n_slots) aslot = 0;
                    if (++bslot == n_bslots) { bslot = 0; bpar ^= 1u; }
                }
                tc_commit(&t_full[st]);
                // input blocks no later tile needs go back to the producers
                uint32_t rslot = first_slot + (max(rel_upto, T.a) - T.a);
                if (rslot >= n_slots) rslot -= n_slots;
                for (uint32_t kb = max(rel_upto, T.a); kb < rel_end; ++kb) {
                    tc_commit(&a_empty[rslot]);
                    if (++rslot == n_slots) rslot = 0;
                }
                RG_TRACE(4, it);
            }
            bslot = __shfl_sync(0xFFFFFFFFu, bslot, 0);
            bpar = __shfl_sync(0xFFFFFFFFu, bpar, 0);
            rel_upto = max(rel_upto, rel_end);
            // slot of the next tile's first block
            if (has_next) {
                if (Tn.a >= staged_upto) first_slot = next_slot;              // gap: its blocks are all new
                else { first_slot += Tn.a - T.a; while (first_slot >= n_slots) first_slot -= n_slots; }
            }
        }
    } else if (warp == RG_WARP_BLOAD) {
        // ================= coefficient-chunk loader =================
        if (lane == 0) {
            uint32_t bslot = 0, bpar = 0;
            RingTileQueue tq;
            tq.init(p, t_begin, t_end);
            for (uint32_t t = t_begin; t < t_end; ++t) {
                const RingTile T = tq.q[0];
                tq.pop(p, t);
                for (uint32_t ks = 0; ks < T.nb; ++ks) {
                    mbar_wait(&b_empty[bslot], bpar ^ 1u);
                    if (p.knockout & 16u) { mbar_arrive(&b_full[bslot]); if (++bslot == p.n_bslots) { bslot = 0; bpar ^= 1u; } continue; }
                    mbar_arrive_expect_tx(&b_full[bslot], TC_B_CHUNK);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(sB + bslot * TC_B_CHUNK)), "l"(p.tile_coef + T.b_off + (uint64_t) ks * TC_B_CHUNK), "r"(TC_B_CHUNK),
                                   "r"(smem_u32(&b_full[bslot])) : "memory");
                    if (++bslot == p.n_bslots) { bslot = 0; bpar ^= 1u; }
                }
            }
        }
    } else {
        // ================= block producers (4 warps, 128 threads) =================
        const uint32_t ptid = tid - RG_WARP_PROD * 32u;
        const uint32_t mg = ptid & 7u, k0 = ptid >> 3;          // this thread stages features k0 and k0 + 16 of a block
        uint32_t slot = 0, par = 0, staged_upto = 0;
        RingTileQueue tq;
        tq.init(p, t_begin, t_end);
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const RingTile T = tq.q[0];
            tq.pop(p, t);
            if (t == t_begin) staged_upto = T.a;
            const uint32_t bt = T.a + T.nb;
            for (uint32_t kb = max(T.a, staged_upto); kb < bt; ++kb) {
                // issue the global loads before waiting for the slot: they do not depend on it
                uint4 w[2][4];
                const uint32_t used_word = kb < p.n_feat_words ? __ldg(p.feat_used + kb) : 0u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t k = k0 + 16u * h;
                    const uint32_t ct = kb * 32u + k;
                    uint32_t sl = NO_SLOT;
                    if (ct < p.n_ct_slots) sl = p.slot_of_ct ? __ldg(p.slot_of_ct + ct) : ct;
                    if (p.knockout & 8u) sl = NO_SLOT;
                    if (sl == NO_SLOT) {
                        if (((used_word >> k) & 1u) && !(p.knockout & 8u)) atomicOr(p.status, 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[h][q] = make_uint4(0, 0, 0, 0);
                    } else {
                        const uint8_t *src = p.in.words + (uint64_t) sl * p.in.stride + 4u * (w_slice + mg * 16u);
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[h][q] = ldg128(src + 16 * q);
                    }
                }
                mbar_wait(&a_empty[slot], par ^ 1u);
                uint8_t *blk = sA + slot * RG_BLOCK_BYTES;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t k = k0 + 16u * h;
                    uint32_t limb[4][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t x0 = __byte_perm(w[h][q].x, w[h][q].y, 0x5140), x1 = __byte_perm(w[h][q].x, w[h][q].y, 0x7362);
                        const uint32_t x2 = __byte_perm(w[h][q].z, w[h][q].w, 0x5140), x3 = __byte_perm(w[h][q].z, w[h][q].w, 0x7362);
                        limb[0][q] = __byte_perm(x0, x2, 0x5410);
                        limb[1][q] = __byte_perm(x0, x2, 0x7632);
                        limb[2][q] = __byte_perm(x1, x3, 0x5410);
                        limb[3][q] = __byte_perm(x1, x3, 0x7632);
                    }
                    uint8_t *dst = blk + (k >> 3) * TC_A_LBO + mg * TC_A_SBO + (k & 7u) * 16u;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4 *>(dst + j * RG_PLANE_BYTES) = make_uint4(limb[j][0], limb[j][1], limb[j][2], limb[j][3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);      // one arrival per producer warp
                if (++slot == p.n_slots) { slot = 0; par ^= 1u; }
            }
            staged_upto = max(staged_upto, bt);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == RG_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
