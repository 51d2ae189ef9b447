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
//   * roles: warps 0-3 epilogue (TMEM lane quadrant = warp id), warp 4 MMA issuer (+ TMEM owner),
//     warp 5 coefficient-image loader (cp.async.bulk), warps 6-9 block producers. mbarrier pipelines:
//     a_full/a_empty per ring slot, b_full/b_empty x2 coefficient buffers, t_full/t_empty x2 TMEM stages
//     (2 x 4 accumulators x 64 columns = all 512 TMEM columns), so the MMAs of tile t+1 overlap the
//     epilogue of tile t and the producers run up to a ring ahead.
#pragma once

#define RG_THREADS 320
#define RG_BLOCK_BYTES (16u * TC_A_LBO)          // 4 planes x 4 feature groups x 1152 B = 18432
#define RG_PLANE_BYTES (4u * TC_A_LBO)
#define RG_MAX_SLOTS 9

struct RingParams {
    const idash_b200_tile *tiles;
    const uint32_t *tile_rows;
    const int32_t *tile_bias;
    const uint8_t *tile_coef;
    const uint32_t *feat_used;     // bit f: feature f is used by some row
    uint32_t n_feat_words;
    uint32_t n_tiles;
    uint32_t n_chunks;             // gridDim.x = 16 * n_chunks
    uint32_t n_slots;              // ring slots (>= widest tile in blocks, + prefetch)
    uint32_t b_buf_bytes;          // size of one coefficient buffer = 2 * tile_kmax * 64
    CtView in, out;
    const uint32_t *slot_of_ct;
    uint32_t n_ct_slots;
    const uint32_t *slot_of_row;
    uint32_t S;
    int *status;
};

__host__ __device__ constexpr uint32_t ring_smem_bytes(uint32_t n_slots, uint32_t kmax) { return n_slots * RG_BLOCK_BYTES + 2u * (2u * kmax * TC_TN); }

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

struct RingTile { uint32_t a, nb; uint64_t b_off; uint32_t flags; };   // first block, blocks, coefficient image

__device__ __forceinline__ RingTile ring_tile(const RingParams &p, uint32_t t) {
    const uint4 *tp = reinterpret_cast<const uint4 *>(p.tiles + t);
    const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    RingTile r;
    r.a = t0.x >> 5; r.nb = t0.y >> 5;
    r.b_off = (uint64_t) t0.z | ((uint64_t) t0.w << 32);
    r.flags = t1.z;
    return r;
}

// Epilogue of one tile for one warp (32 words x 64 rows). FAST: rows are consecutive output slots, the row
// stride is a compile-time constant, so every store address is base + immediate.
template <bool FAST, uint32_t STRIDE, bool BIAS>
__device__ __forceinline__ void ring_epilogue(uint32_t tmem_lane_col, uint8_t *base, uint64_t ptr_lo, uint64_t ptr_hi,
                                              uint32_t bias_lo, uint32_t bias_hi, uint32_t bias_flag, uint32_t lane_off) {
#pragma unroll
    for (uint32_t chunk = 0; chunk < 8; ++chunk) {
        const uint32_t col0 = chunk * 8u;
        uint32_t v0[8], v1[8], v2[8], v3[8];
        tc_ld8(tmem_lane_col + col0, v0);
        tc_ld8(tmem_lane_col + col0 + TC_TN, v1);
        tc_ld8(tmem_lane_col + col0 + 2 * TC_TN, v2);
        tc_ld8(tmem_lane_col + col0 + 3 * TC_TN, v3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const uint32_t n = col0 + c;
            const uint32_t src = n & 31u;
            uint32_t v = ((v3[c] * 256u + v2[c]) * 256u + v1[c]) * 256u + v0[c];
            if (BIAS) v += __shfl_sync(0xFFFFFFFFu, n < 32 ? bias_lo : bias_hi, src) * bias_flag;
            if (FAST) {
                stg32_stream(base + (uint64_t) n * STRIDE + lane_off, v);
            } else {
                const uint64_t ptr = __shfl_sync(0xFFFFFFFFu, n < 32 ? ptr_lo : ptr_hi, src);
                if (ptr) stg32_stream(reinterpret_cast<uint8_t *>(ptr) + lane_off, v);
            }
        }
    }
}

__global__ void __launch_bounds__(RG_THREADS, 1) cloud_ring_kernel(const RingParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[RG_MAX_SLOTS], a_empty[RG_MAX_SLOTS], b_full[2], b_empty[2], t_full[2], t_empty[2];
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
        for (uint32_t s = 0; s < p.n_slots; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
        for (uint32_t s = 0; s < 2; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp < 4) {
        // ================= epilogue =================
        const uint32_t word_in_slice = warp * 32u + lane;
        const uint32_t lane_off = 4u * word_in_slice;
        const uint32_t bias_flag = (is_b && i_slice + word_in_slice < p.S) ? 1u : 0u;
        const bool records = p.out.records != 0;
        uint32_t it = 0;
        // row information of the next tile is fetched one tile ahead
        uint32_t row_lo = __ldg(p.tile_rows + (uint64_t) t_begin * TC_TN + lane), row_hi = __ldg(p.tile_rows + (uint64_t) t_begin * TC_TN + 32 + lane);
        int32_t b_lo = __ldg(p.tile_bias + (uint64_t) t_begin * TC_TN + lane), b_hi = __ldg(p.tile_bias + (uint64_t) t_begin * TC_TN + 32 + lane);
        uint32_t flags = __ldg(&p.tiles[t_begin].flags);
        for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
            const uint32_t cur_lo = row_lo, cur_hi = row_hi, cur_flags = flags;
            const uint32_t bias_lo = (uint32_t) b_lo * (uint32_t) IDASH_B200_ONE_IN_T32, bias_hi = (uint32_t) b_hi * (uint32_t) IDASH_B200_ONE_IN_T32;
            if (t + 1 < t_end) {
                row_lo = __ldg(p.tile_rows + (uint64_t) (t + 1) * TC_TN + lane); row_hi = __ldg(p.tile_rows + (uint64_t) (t + 1) * TC_TN + 32 + lane);
                b_lo = __ldg(p.tile_bias + (uint64_t) (t + 1) * TC_TN + lane); b_hi = __ldg(p.tile_bias + (uint64_t) (t + 1) * TC_TN + 32 + lane);
                flags = __ldg(&p.tiles[t + 1].flags);
            }
            const bool fast = (cur_flags & 1u) && p.slot_of_row == nullptr;
            uint64_t ptr_lo = 0, ptr_hi = 0;
            uint8_t *base = nullptr;
            if (fast) {
                const uint32_t row0 = __shfl_sync(0xFFFFFFFFu, cur_lo, 0);
                base = p.out.words + (uint64_t) row0 * p.out.stride + 4u * w_slice;
            } else {
                if (cur_lo != IDASH_B200_NO_ROW) ptr_lo = (uint64_t) (p.out.words + (uint64_t) (p.slot_of_row ? __ldg(p.slot_of_row + cur_lo) : cur_lo) * p.out.stride + 4u * w_slice);
                if (cur_hi != IDASH_B200_NO_ROW) ptr_hi = (uint64_t) (p.out.words + (uint64_t) (p.slot_of_row ? __ldg(p.slot_of_row + cur_hi) : cur_hi) * p.out.stride + 4u * w_slice);
            }
            const uint32_t st = it & 1u;
            mbar_wait(&t_full[st], (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem + ((warp * 32u) << 16) + st * 4u * TC_TN;
            if (fast) {
                if (records) {
                    if (is_b) ring_epilogue<true, IDASH_B200_RECORD_BYTES, true>(tl, base, 0, 0, bias_lo, bias_hi, bias_flag, lane_off);
                    else ring_epilogue<true, IDASH_B200_RECORD_BYTES, false>(tl, base, 0, 0, 0, 0, 0, lane_off);
                } else {
                    if (is_b) ring_epilogue<true, IDASH_B200_CT_BYTES, true>(tl, base, 0, 0, bias_lo, bias_hi, bias_flag, lane_off);
                    else ring_epilogue<true, IDASH_B200_CT_BYTES, false>(tl, base, 0, 0, 0, 0, 0, lane_off);
                }
            } else {
                ring_epilogue<false, 0, true>(tl, nullptr, ptr_lo, ptr_hi, bias_lo, bias_hi, bias_flag, lane_off);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&t_empty[st]);
        }
    } else if (warp == 4) {
        // ================= MMA issuer =================
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint32_t seq_end = 0, staged_upto = 0, rel_upto = 0, it = 0;
        RingTile T = ring_tile(p, t_begin);
        staged_upto = rel_upto = T.a;
        for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
            RingTile Tn = T;
            const bool has_next = t + 1 < t_end;
            if (has_next) Tn = ring_tile(p, t + 1);
            const uint32_t bt = T.a + T.nb;
            const uint32_t first_new = max(T.a, staged_upto);
            const uint32_t seq_before = seq_end;
            seq_end += bt > first_new ? bt - first_new : 0u;
            staged_upto = max(staged_upto, bt);
            const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(&t_empty[st], ph ^ 1u);
            mbar_wait(&b_full[st], ph);
            for (uint32_t s = seq_before; s < seq_end; ++s) mbar_wait(&a_full[s % p.n_slots], (s / p.n_slots) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t K = T.nb * 32u;
                const uint32_t d0 = tmem + st * 4u * TC_TN;
                for (uint32_t ks = 0; ks < T.nb; ++ks) {
                    const uint32_t sq = seq_end - (T.nb - ks);          // sequence number of block a + ks
                    const uint32_t ablk = a0 + (sq % p.n_slots) * RG_BLOCK_BYTES;
                    const uint32_t bblk = b0 + st * p.b_buf_bytes + ks * 2u * TC_B_LBO;
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {
#pragma unroll
                        for (uint32_t i = 0; i < 2; ++i) {
                            if (i + j > 3) continue;
                            const uint64_t da = tc_desc(ablk + j * RG_PLANE_BYTES, TC_A_LBO, TC_A_SBO);
                            const uint64_t db = tc_desc(bblk + i * K * TC_TN, TC_B_LBO, TC_B_SBO);
                            const uint32_t first = (ks == 0) && (i == 1 || j == 0);
                            tc_mma(d0 + (i + j) * TC_TN, da, db, tc_idesc(i), first ? 0u : 1u);
                        }
                    }
                }
                tc_commit(&t_full[st]);
                tc_commit(&b_empty[st]);
                // blocks no later tile needs go back to the producers
                const uint32_t rel_end = has_next ? min(Tn.a, bt) : bt;
                for (uint32_t kb = max(rel_upto, T.a); kb < rel_end; ++kb) tc_commit(&a_empty[(seq_end - (bt - kb)) % p.n_slots]);
                rel_upto = max(rel_upto, rel_end);
            } else {
                const uint32_t rel_end = has_next ? min(Tn.a, bt) : bt;
                rel_upto = max(rel_upto, rel_end);
            }
            __syncwarp();
            T = Tn;
        }
    } else if (warp == 5) {
        // ================= coefficient-image loader =================
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
                const RingTile T = ring_tile(p, t);
                const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
                const uint32_t bytes = 2u * T.nb * 32u * TC_TN;
                mbar_wait(&b_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&b_full[st], bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sB + st * p.b_buf_bytes)), "l"(p.tile_coef + T.b_off), "r"(bytes), "r"(smem_u32(&b_full[st])) : "memory");
            }
        }
    } else {
        // ================= block producers (warps 6-9, 128 threads) =================
        const uint32_t ptid = tid - 6u * 32u;
        const uint32_t mg = ptid & 7u, k0 = ptid >> 3;          // this thread stages features k0 and k0 + 16 of a block
        uint32_t seq = 0, staged_upto = 0;
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const RingTile T = ring_tile(p, t);
            if (t == t_begin) staged_upto = T.a;
            const uint32_t bt = T.a + T.nb;
            for (uint32_t kb = max(T.a, staged_upto); kb < bt; ++kb, ++seq) {
                const uint32_t slot = seq % p.n_slots;
                // issue the global loads before waiting for the slot: they do not depend on it
                uint4 w[2][4];
                const uint32_t used_word = kb < p.n_feat_words ? __ldg(p.feat_used + kb) : 0u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t k = k0 + 16u * h;
                    const uint32_t ct = kb * 32u + k;
                    uint32_t sl = NO_SLOT;
                    if (ct < p.n_ct_slots) sl = p.slot_of_ct ? __ldg(p.slot_of_ct + ct) : ct;
                    if (sl == NO_SLOT) {
                        if ((used_word >> k) & 1u) atomicOr(p.status, 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[h][q] = make_uint4(0, 0, 0, 0);
                    } else {
                        const uint8_t *src = p.in.words + (uint64_t) sl * p.in.stride + 4u * (w_slice + mg * 16u);
#pragma unroll
                        for (int q = 0; q < 4; ++q) w[h][q] = ldg128(src + 16 * q);
                    }
                }
                mbar_wait(&a_empty[slot], ((seq / p.n_slots) & 1u) ^ 1u);
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
                mbar_arrive(&a_full[slot]);
            }
            staged_upto = max(staged_upto, bt);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
