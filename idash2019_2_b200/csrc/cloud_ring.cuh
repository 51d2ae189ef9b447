// K2r: persistent, warp-specialised variant of the tensor-core cloud kernel (any NUM_REGIONS: rotations are applied
// while an input block is staged). Same arithmetic as cloud_tc.cuh; what changes is the schedule:
//
//   * grid = 16 slices x C chunks (C = SMs / 16): CTA (slice, chunk) walks the tiles of its chunk in order
//     for ONE 128-word slice of the ciphertext axis. The 16 CTAs of a chunk advance together, so whole
//     8 KB ciphertexts are read / written at about the same time and a tile's coefficient image is
//     fetched from HBM once and served to the other 15 CTAs by L2. The SMs that division leaves over (148 = 16 x 9 + 4)
//     run one more, short chunk in waves (RingParams::extra_tiles): the kernel's time follows the number of SMs it uses.
//   * the limb planes of the input blocks (32 features x 128 words x 4 planes = 18 KB) live in a
//     shared-memory RING: consecutive tiles share most of their band, so every (block, slice) is loaded
//     from global memory, split into byte planes and stored ONCE per CTA -- the north-star's "stage each
//     overlapping tag window once and reuse it across consecutive target SNPs".
//   * roles: warps 0-7 epilogue (TMEM lane quadrant = warp id % 4, rows 32 (warp id / 4) .. +31 of the tile),
//     warps 8 and 15 MMA issuers (even / odd tiles = TMEM stage 0 / 1; warp 8 owns the TMEM allocation), warp 9 coefficient
//     loader (cp.async.bulk copies of a tile's image into a ring of 4 KB chunks, of its metadata record into a ring of 8),
//     warps 10-13 block producers, warp 14 progress publisher.
//   * the MMA warps wait for input blocks by PARITY on reused ring slots; two warps on alternating tiles can be a whole use
//     early where one warp could not, hence the gate on waits_done_s (see the MMA issuers).
//   * synchronisation is built so that the MMA warp -- the one serial instruction stream every tile passes
//     through -- does as little as possible per tile: its waits (TMEM stage free, coefficient stage full, new
//     input blocks full) are taken by different lanes at the same time, and it issues ONE tcgen05.commit per
//     tile (t_full). Everything that has to know "the MMAs of tile t are complete" (the coefficient loader and
//     the block producers, to reuse ring space) reads two monotonic shared-memory counters that the
//     publisher warp advances after it has seen t_full complete. (Measured: a tcgen05.commit costs the
//     issuing thread ~150 cycles and an mbarrier wait ~120 even when satisfied; with per-slot commits and
//     serial waits the MMA warp, not HBM, set the tile rate.)
//   * TMEM: 2 stages x 4 accumulators x 64 columns = all 512 columns, so the MMAs of tile t+1 overlap the
//     epilogue of tile t.
//   * batched launches (template parameter BATCHED): the same model on several input / output sets -- what a GPU that
//     owns a target range does for several sample batches. The tile list is walked once per set as VIRTUAL tiles;
//     see RingParams::n_batches. One launch instead of N pays the ramp-up and the tail of the persistent grid once.
#pragma once

#define RG_THREADS 512
#define RG_EPI_WARPS 8
#define RG_WARP_MMA 8
#define RG_WARP_BLOAD 9
#define RG_WARP_PROD 10
#define RG_WARP_PUB 14
#define RG_WARP_MMA2 15
#define RG_BLOCK_BYTES (16u * TC_A_LBO)          // 4 planes x 4 feature groups x 1152 B = 18432
#define RG_PLANE_BYTES (4u * TC_A_LBO)
#define RG_MAX_SLOTS 9
#define RG_BBARS 8u                              // per-tile "coefficients landed" barriers (tile it uses b_full[it % 8])
#define RG_META_STAGES RG_BBARS                  // per-tile metadata ring: tile it uses record it % 8 (caller rows + Constants, 512 bytes)
#define RG_META_BYTES (RG_META_STAGES * 8u * TC_TN)
#define RG_MAX_BATCHES 8u
#define RG_BATCH_SHIFT 14u                       // input blocks per batch < 2^14 (features < 524 288) in a batched launch
#define RG_SMEM_MAX (226u * 1024u)               // dynamic shared memory budget of the one resident CTA
// Production schedule (measured on B200, profiles/r02_ring_schedule_sweep.txt): warp-converged MMA issue + TMEM stage handed back
// after the epilogue's last tcgen05.ld (tune 8 | 64) + x16 TMEM loads (EPI 1): 0.470 / 0.507 / 0.602 / 0.537 ms at neighbors 5 / 20 /
// 50 / NUM_REGIONS = 3, against 0.480 / 0.514 / 0.634 / 0.552 ms without the last two.
#define RG_TUNE_DEFAULT (8u | 64u | 128u)        // RingParams::tune of the production library: NUM_REGIONS = 1 ...
#define RG_TUNE_ROT (8u | 64u)                   // ... and NUM_REGIONS > 1 (weight-stationary MMAs measured slower there: 0.58 vs 0.54 ms)
#define RG_EPI_DEFAULT 1                         // ... and its epilogue schedule (template parameter EPI of the kernel)

struct RingParams {
    const idash_b200_tile *tiles;
    const uint32_t *tile_rows;
    const int32_t *tile_bias;
    const uint8_t *tile_coef;
    const uint32_t *feat_used;     // bit f: feature f is used by some row
    uint32_t n_feat_words;
    uint32_t n_tiles;              // tiles of this launch: [tile_base, tile_base + n_tiles)
    uint32_t tile_base;
    uint32_t n_chunks;             // regular chunks; gridDim.x = n_slices * (n_chunks + (extra_tiles ? 1 : 0))
    uint32_t extra_tiles;          // virtual tiles of the short extra chunk at the end of the list (0: none), see the kernel
    uint32_t n_slices;             // 128-word slices of the ciphertext axis that are COMPUTED: 16, or 8 + ceil(RS / 128) when
                                   // NUM_REGIONS > 1 (b[RS..N) is zero, eval/idash.cpp:839-841: those slices are only zero-filled)
    uint32_t n_slots;              // input-block ring slots (>= widest tile in blocks, + prefetch)
    uint32_t n_bchunks;            // coefficient ring: 4096-byte chunks (one 32-feature K step each), a tile takes K / 32 of them
    uint64_t coef_bytes;           // size of the coefficient image array (bound for the L2 prefetch)
    uint32_t coef_prefetch;        // the loader prefetches the images this many tiles ahead into L2 (0 = off)
    uint32_t meta_off;             // byte offset of the ring of per-tile metadata records (RG_META_STAGES x {rows[64], Constant[64]})
    uint32_t zero_off;             // NUM_REGIONS > 1: byte offset of RG_ZERO_BYTES of zeros (512-byte aligned)
    uint32_t hdr_off;              // byte offset (dynamic shared memory) of the CTA's tile-header table
    uint32_t max_chunk_tiles;      // capacity of that table (tiles per chunk, rounded up)
    CtView in, out;
    // Batched launch (n_batches > 1): the same tiles are evaluated for several input / output sets that share the model. The CTAs
    // walk VIRTUAL tiles v = batch * n_tiles + tile; a batch's input blocks get the virtual index batch << RG_BATCH_SHIFT | block, so
    // the band still only moves forward (a batch boundary looks like a gap) and no role needs to know about batches except where
    // it forms a global address. in / out describe batch 0 (layout, stride, count are common to all batches).
    uint32_t n_batches;
    unsigned long long batch_in[RG_MAX_BATCHES], batch_out[RG_MAX_BATCHES];   // words pointers of every batch
    // ... and several MODELS: batch b may bring its own model of the same shape (same row count, hence tile count, same NUM_REGIONS /
    // REGION_SIZE; NUM_SAMPLES and all coefficients may differ) -- BASELINE configs[3], the population-stratified model sets evaluated
    // in one launch. For batches that share the model the entries repeat tiles / tile_rows / ... above.
    const idash_b200_tile *batch_tiles[RG_MAX_BATCHES];
    const uint32_t *batch_rows[RG_MAX_BATCHES];
    const int32_t *batch_bias[RG_MAX_BATCHES];
    const uint8_t *batch_coef[RG_MAX_BATCHES];
    const uint32_t *batch_feat_used[RG_MAX_BATCHES];
    uint32_t batch_nfw[RG_MAX_BATCHES], batch_S[RG_MAX_BATCHES];
    unsigned long long batch_coef_bytes[RG_MAX_BATCHES];
    const uint32_t *slot_of_ct;
    uint32_t n_ct_slots;
    const uint32_t *slot_of_row;
    uint32_t S, NR, RS;            // NUM_SAMPLES, NUM_REGIONS, REGION_SIZE (feature f = ciphertext f / NR rotated by (f % NR) * RS words)
    int *status;
    uint32_t trace_cta;            // 0 = off, else 1 + index of the CTA whose timeline is recorded
    uint32_t tune;                 // schedule switches: 1 epilogue waits with try_wait, 2 publisher waits with try_wait, 4 MMA warp polls
                                   // without nanosleep, 8 warp-converged MMA issue (uniform operands), 32 the two MMA warps issue their
                                   // tiles strictly in tile order, 64 the epilogue frees its TMEM stage after its last tcgen05.ld, 128 weight-stationary MMAs (the
                                   // coefficient chunk of a K step is read from shared memory once, tcgen05.mma.ws + collector buffer); host side: 256 / 512 pick
                                   // the kernel's EPI template parameter (x16 loads / burst epilogue)
    uint32_t knockout;             // profiling aid (IDASH_B200_KNOCKOUT, results are wrong when non-zero):
                                   // 1 no MMAs, 2 no output stores, 4 no epilogue TMEM loads, 8 no input loads, 16 no zero fill, 32 no bias / mask,
                                   // 64 producers sleep 20 us before every block (results stay right: tests/test_gpu_cloud.py), 128 without the gate of the
                                   // MMA warps' slot waits (with 64: the stale-slot race the gate closes shows)
};

#define RG_ZERO_BYTES 4096u                      // NUM_REGIONS > 1: a block of zeros, the source of the bulk stores that fill b[RS'..N) of every output row
__host__ __device__ constexpr uint32_t ring_smem_bytes(uint32_t n_slots, uint32_t n_bchunks, uint32_t max_chunk_tiles, bool rot) {
    return n_slots * RG_BLOCK_BYTES + n_bchunks * TC_B_CHUNK + RG_META_BYTES + (rot ? RG_ZERO_BYTES : 0u) + 4u * max_chunk_tiles;
}

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
// non-blocking probe (no hardware suspend): 1 if the phase with this parity has completed
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar_addr, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void mbar_spin(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    while (!mbar_test(a, parity)) __nanosleep(40);
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
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred;
}
// tcgen05.mma issued by the elected lane of a converged warp; the operands are warp-uniform
__device__ __forceinline__ void tc_mma_p(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tc_mma_kstep_p(uint32_t d0, uint64_t da, uint32_t plane_units, uint64_t db, bool first, uint32_t leader) {
    const uint32_t acc = first ? 0u : 1u;
    tc_mma_p(d0 + 0 * TC_TN, da + 0 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc, leader);
    tc_mma_p(d0 + 2 * TC_TN, da + 2 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc, leader);
    tc_mma_p(d0 + 1 * TC_TN, da + 1 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), 1u, leader);
    tc_mma_p(d0 + 3 * TC_TN, da + 3 * (uint64_t) plane_units, db, tc_idesc(TC_TN), 1u, leader);
}
// Weight-stationary form (tune & 128): the four MMAs of a K step multiply four different limb planes by the SAME coefficient chunk.
// tcgen05.mma.ws keeps the B operand in a collector buffer of the tensor core (fill -> use -> lastuse), so the chunk is read from
// shared memory once per K step instead of three times: an N = 128, K = 32 i8 MMA reads 8 KB of operands in its 64 cycles -- all
// of the 128 B/clk the shared-memory pipe has -- and every other shared-memory access of the CTA (TMA writes of the coefficient
// images, the producers' operand stores) slows the MMAs down (measured in the kernel: 84 cycles per MMA instead of 64).
// BUF: collector buffer of the issuing warp (the two MMA warps interleave in the pipe and must not share one).
#define TC_MMA_WS(BUFSTR, USAGE)                                                                                            \
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"                               \
                 "@q tcgen05.mma.ws.cta_group::1.kind::i8.collector::" BUFSTR "::" USAGE " [%0], %1, %2, %3, p;\n\t}\n"          \
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(leader) : "memory")
template <int BUF, int USAGE>   // USAGE 0 fill, 1 use, 2 lastuse, 3 discard
__device__ __forceinline__ void tc_mma_ws_p(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    if (BUF == 0) {
        if (USAGE == 0) TC_MMA_WS("b0", "fill"); else if (USAGE == 1) TC_MMA_WS("b0", "use");
        else if (USAGE == 2) TC_MMA_WS("b0", "lastuse"); else TC_MMA_WS("b0", "discard");
    } else {
        if (USAGE == 0) TC_MMA_WS("b1", "fill"); else if (USAGE == 1) TC_MMA_WS("b1", "use");
        else if (USAGE == 2) TC_MMA_WS("b1", "lastuse"); else TC_MMA_WS("b1", "discard");
    }
}
template <int BUF>
__device__ __forceinline__ void tc_mma_kstep_ws_p(uint32_t d0, uint64_t da, uint32_t plane_units, uint64_t db, bool first, uint32_t leader) {
    const uint32_t acc = first ? 0u : 1u;
    tc_mma_ws_p<BUF, 0>(d0 + 0 * TC_TN, da + 0 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc, leader);
    tc_mma_ws_p<BUF, 1>(d0 + 2 * TC_TN, da + 2 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), acc, leader);
    tc_mma_ws_p<BUF, 2>(d0 + 1 * TC_TN, da + 1 * (uint64_t) plane_units, db, tc_idesc(2 * TC_TN), 1u, leader);
    tc_mma_ws_p<BUF, 3>(d0 + 3 * TC_TN, da + 3 * (uint64_t) plane_units, db, tc_idesc(TC_TN), 1u, leader);
}
// monotonic progress counters in shared memory
__device__ __forceinline__ void progress_publish(uint32_t *ctr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(ctr)), "r"(v) : "memory");
}
// ... advanced by several threads in no fixed order (release: what the publishing warp has observed is visible to whoever reads the value)
__device__ __forceinline__ void progress_publish_max(uint32_t *ctr, uint32_t v) {
    asm volatile("fence.acq_rel.cta;\n\tred.relaxed.cta.shared::cta.max.u32 [%0], %1;" ::"r"(smem_u32(ctr)), "r"(v) : "memory");
}
__device__ __forceinline__ void progress_wait(const uint32_t *ctr, uint32_t at_least) {
    uint32_t v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(ctr)) : "memory");
        if ((int32_t) (v - at_least) >= 0) break;
        __nanosleep(64);      // the waiters are far ahead of the data path: do not spin hot next to the MMA warp
    }
}

// tracing aid (IDASH_B200_TRACE=<cta>): per-tile SM-clock timestamps of one CTA, see tools/trace_ring.py
#define RG_TRACE_TILES 96
#define RG_TRACE_EVENTS 12
__device__ unsigned long long g_ring_trace[RG_TRACE_TILES * RG_TRACE_EVENTS];
#define RG_TRACE(ev, it_)                                                                                     \
    do {                                                                                                      \
        if (k_trace == blockIdx.x + 1u && (it_) >= 64u && (it_) < 64u + RG_TRACE_TILES)                     \
            g_ring_trace[((it_) - 64u) * RG_TRACE_EVENTS + (ev)] = clock64();                                    \
    } while (0)

struct RingTile { uint32_t a, nb, fast; };   // first input block, blocks, rows are consecutive caller rows

// Tile headers: every role walks the tiles of the chunk, and a header read from global memory inside those loops
// exposes an L2 / HBM round trip per tile per role (measured: 40% of the epilogue warps' time). The CTA therefore
// packs the headers of its chunk into shared memory once, before the roles start:
//     bits 0..19 first block (f_base / 32), bits 20..30 blocks (K / 32), bit 31 flags & 1
#define RG_HDR_A_BITS 20
__device__ __forceinline__ uint32_t ring_pack_hdr(const idash_b200_tile *tp) {
    const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(tp));
    const uint32_t flags = __ldg(&tp->flags);
    return (t0.x >> 5) | ((t0.y >> 5) << RG_HDR_A_BITS) | ((flags & 1u) << 31);
}
__device__ __forceinline__ RingTile ring_hdr(const uint32_t *hdr_s, uint32_t i) {
    const uint32_t h = hdr_s[i];
    RingTile r;
    r.a = h & ((1u << RG_HDR_A_BITS) - 1u); r.nb = (h >> RG_HDR_A_BITS) & 0x7FFu; r.fast = h >> 31;
    return r;
}

#define RG_TRACE_V(ev, it_, val)                                                                             \
    do {                                                                                                      \
        if (k_trace == blockIdx.x + 1u && (it_) >= 64u && (it_) < 64u + RG_TRACE_TILES)                     \
            g_ring_trace[((it_) - 64u) * RG_TRACE_EVENTS + (ev)] = (unsigned long long) (val);                    \
    } while (0)

// The walk over the tiles of a chunk that every role repeats identically: which input blocks a tile adds to the
// ring (staged in order, one slot each) and which it lets go of once its MMAs are complete.
struct RingWalk {
    uint32_t staged_upto, rel_upto;   // blocks < staged_upto have been staged, blocks < rel_upto released
    __device__ __forceinline__ void init(uint32_t a0) { staged_upto = rel_upto = a0; }
    // new blocks of tile T: [first_new, T.a + T.nb)
    __device__ __forceinline__ uint32_t first_new(const RingTile &T) const { return max(T.a, staged_upto); }
    // blocks tile T releases, given the next tile (or none): [rel_begin, rel_end)
    __device__ __forceinline__ void advance(const RingTile &T, const RingTile &Tn, bool has_next, uint32_t &rel_begin, uint32_t &rel_end) {
        const uint32_t bt = T.a + T.nb;
        staged_upto = max(staged_upto, bt);
        rel_begin = max(rel_upto, T.a);
        rel_end = has_next ? min(Tn.a, bt) : bt;
        rel_end = max(rel_end, rel_begin);
        rel_upto = max(rel_upto, rel_end);
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
// release_bar: the TMEM stage's t_empty barrier when the stage is to be handed back as soon as the last tcgen05.ld of this warp
// has completed (the accumulators are in registers then; the recombine + stores of the last group no longer need TMEM), or
// nullptr when the caller arrives after the whole epilogue. Returns true if it has arrived.
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// Constants of four consecutive tile rows from the tile's metadata record in shared memory (one broadcast 128-bit load; the record is
// brought in by the coefficient loader's bulk copies, so the epilogue issues no global load that could queue behind its own stores)
template <bool BIAS>
__device__ __forceinline__ uint4 ring_bias4(uint32_t bias_addr, uint32_t n) {
    uint4 q = make_uint4(0, 0, 0, 0);
    if (BIAS) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(bias_addr + 4u * n));
    return q;
}
__device__ __forceinline__ uint32_t ring_q(const uint4 &q, uint32_t i) { return i == 0 ? q.x : i == 1 ? q.y : i == 2 ? q.z : q.w; }

// One recombined output word -> its row (shared by the two load widths below)
template <bool FAST, uint32_t STRIDE, bool BIAS, bool MASK>
__device__ __forceinline__ void ring_emit(uint32_t n, uint32_t x, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_n, uint32_t bias_flag,
                                          uint32_t lane_off, uint32_t knockout, uint32_t keep_mask) {
    if (BIAS) x += bias_n * bias_flag;          // bias_flag = 2^18 (ONE_IN_T32) on the words b[0..S), else 0 (idash.cpp:805-810)
    if (MASK) x &= keep_mask;                                                  // b[RS..N) = 0 (idash.cpp:839-841)
    if (knockout & 2u) { if (x == 0x9E3779B9u && bias_flag == 77u) stg32_stream(base_lane, x); return; }
    if (FAST) {
        stg32_stream(base_lane + (uint64_t) n * STRIDE, x);
    } else {
        const uint64_t ptr = __shfl_sync(0xFFFFFFFFu, ptr_own, n);
        if (ptr) stg32_stream(reinterpret_cast<uint8_t *>(ptr) + lane_off, x);
    }
}

// Wide-load variant (tune & 256): two groups of 16 rows, each accumulator read with ONE 32x32b.x16 load -- 8 tcgen05.ld per
// warp and tile instead of 16. TMEM is shared by the epilogue's loads and the accumulator read-modify-writes of the running
// MMAs (knock-outs: MMAs alone 0.42 ms, TMEM loads + recombine alone 0.41-0.51 ms, together 0.63 ms at neighbors = 50).
template <bool FAST, uint32_t STRIDE, bool BIAS, bool MASK>
__device__ __forceinline__ bool ring_epilogue_w16(uint32_t taddr, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_own, uint32_t bias_flag,
                                                  uint32_t lane_off, uint32_t knockout, uint32_t keep_mask, uint64_t *release_bar) {
    if (knockout & 4u) return false;
    uint32_t v[4][16];
#pragma unroll
    for (uint32_t g = 0; g < 2; ++g) {
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) tc_ld16(taddr + g * 16u + j * TC_TN, v[j]);
        // the group's 16 Constants come in while the TMEM loads are in flight (four broadcast 128-bit loads; on wide bands the MMAs'
        // operand fetch saturates the shared-memory pipe and a load issued right before its use stalls the stores behind it)
        uint4 bq[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) bq[k] = ring_bias4<BIAS>(bias_own, g * 16u + 4u * k);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (g == 1 && release_bar) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive(release_bar);
        }
#pragma unroll
        for (uint32_t c = 0; c < 16; ++c) {
            const uint32_t x = ((v[3][c] * 256u + v[2][c]) * 256u + v[1][c]) * 256u + v[0][c];
            ring_emit<FAST, STRIDE, BIAS, MASK>(g * 16u + c, x, base_lane, ptr_own, ring_q(bq[c >> 2], c & 3u), bias_flag, lane_off, knockout, keep_mask);
        }
    }
    return release_bar != nullptr;
}

// Packed burst variant (EPI = 2): ALL accumulators of the warp's 32 rows are pulled out of TMEM with ONE round of loads, the
// stage is handed back, and only then are the 32 rows recombined and stored. With the two-group schedule above the stage is
// released after the first 16 rows have been stored -- under HBM back-pressure ~1000+ cycles after the tile completed -- and on
// wide bands, where the tensor pipe is as busy as HBM, the MMAs of tile t + 2 wait for exactly that (2 TMEM stages; trace:
// profiles/r02_trace_ring_n50.txt). Registers: out = P0 + 2^8 P1 + 2^16 P2 + 2^24 P3 mod 2^32 needs all of P0, 24 bits of P1, 16 of P2
// and 8 of P3, so P2 and P3 are read with .pack::16b (the low halves of two adjacent columns in one register): 32 + 32 + 16 + 16 =
// 96 data registers instead of 128.
__device__ __forceinline__ void tc_ld16_pack(uint32_t taddr, uint32_t (&v)[16]) {     // 32 columns, low 16 bits each
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
template <bool FAST, uint32_t STRIDE, bool BIAS, bool MASK>
__device__ __forceinline__ bool ring_epilogue_burst(uint32_t taddr, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_own, uint32_t bias_flag,
                                                    uint32_t lane_off, uint32_t knockout, uint32_t keep_mask, uint64_t *release_bar) {
    if (knockout & 4u) return false;
    uint32_t p0[2][16], p1[2][16], p2[16], p3[16];
    tc_ld16(taddr, p0[0]);
    tc_ld16(taddr + 16u, p0[1]);
    tc_ld16(taddr + TC_TN, p1[0]);
    tc_ld16(taddr + TC_TN + 16u, p1[1]);
    tc_ld16_pack(taddr + 2 * TC_TN, p2);
    tc_ld16_pack(taddr + 3 * TC_TN, p3);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (release_bar) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31u) == 0) mbar_arrive(release_bar);
    }
#pragma unroll
    for (uint32_t n4 = 0; n4 < 32; n4 += 4) {
        const uint4 bq = ring_bias4<BIAS>(bias_own, n4);
#pragma unroll
        for (uint32_t n = n4; n < n4 + 4; ++n) {
            const uint32_t q2 = p2[n >> 1], q3 = p3[n >> 1];
            // column n of P2 / P3: low half of the pair register for even n, high half for odd n
            const uint32_t hi = (n & 1u) ? (q2 & 0xFFFF0000u) + ((q3 & 0x00FF0000u) << 8) : (q2 << 16) + (q3 << 24);
            const uint32_t x = p0[n >> 4][n & 15u] + (p1[n >> 4][n & 15u] << 8) + hi;
            ring_emit<FAST, STRIDE, BIAS, MASK>(n, x, base_lane, ptr_own, ring_q(bq, n - n4), bias_flag, lane_off, knockout, keep_mask);
        }
    }
    return release_bar != nullptr;
}

// Overlapped x16 variant (EPI = 3): as EPI = 1, but the loads of the second group of 16 rows are issued BEFORE the first group is
// stored (the first group is recombined into 16 registers first, so 16 + 64 data registers are live): the TMEM round trip of the
// second group -- exposed in EPI = 1, ~150-300 cycles per tile in which the warp issues nothing -- hides behind the 16 stores.
__device__ __forceinline__ void tc_ld8_pack(uint32_t taddr, uint32_t (&v)[8]) {     // 16 columns, low 16 bits each
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
template <bool FAST, uint32_t STRIDE, bool BIAS, bool MASK>
__device__ __forceinline__ bool ring_epilogue_w16o(uint32_t taddr, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_own, uint32_t bias_flag,
                                                   uint32_t lane_off, uint32_t knockout, uint32_t keep_mask, uint64_t *release_bar) {
    if (knockout & 4u) return false;
    uint32_t x0[16];
    {
        uint32_t v[4][16];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) tc_ld16(taddr + j * TC_TN, v[j]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (uint32_t c = 0; c < 16; ++c) x0[c] = ((v[3][c] * 256u + v[2][c]) * 256u + v[1][c]) * 256u + v[0][c];
    }
    // second group: P2 / P3 packed (16 + 16 + 8 + 8 = 48 registers beside the 16 recombined words of the first group)
    uint32_t p0[16], p1[16], p2[8], p3[8];
    tc_ld16(taddr + 16u, p0);
    tc_ld16(taddr + 16u + TC_TN, p1);
    tc_ld8_pack(taddr + 16u + 2 * TC_TN, p2);
    tc_ld8_pack(taddr + 16u + 3 * TC_TN, p3);
#pragma unroll
    for (uint32_t c4 = 0; c4 < 16; c4 += 4) {
        const uint4 bq = ring_bias4<BIAS>(bias_own, c4);
#pragma unroll
        for (uint32_t c = c4; c < c4 + 4; ++c)
            ring_emit<FAST, STRIDE, BIAS, MASK>(c, x0[c], base_lane, ptr_own, ring_q(bq, c - c4), bias_flag, lane_off, knockout, keep_mask);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (release_bar) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31u) == 0) mbar_arrive(release_bar);
    }
#pragma unroll
    for (uint32_t c4 = 0; c4 < 16; c4 += 4) {
        const uint4 bq = ring_bias4<BIAS>(bias_own, 16u + c4);
#pragma unroll
        for (uint32_t c = c4; c < c4 + 4; ++c) {
            const uint32_t q2 = p2[c >> 1], q3 = p3[c >> 1];
            const uint32_t hi = (c & 1u) ? (q2 & 0xFFFF0000u) + ((q3 & 0x00FF0000u) << 8) : (q2 << 16) + (q3 << 24);
            const uint32_t x = p0[c] + (p1[c] << 8) + hi;
            ring_emit<FAST, STRIDE, BIAS, MASK>(16u + c, x, base_lane, ptr_own, ring_q(bq, c - c4), bias_flag, lane_off, knockout, keep_mask);
        }
    }
    return release_bar != nullptr;
}

template <int EPI, bool FAST, uint32_t STRIDE, bool BIAS, bool MASK = false>
__device__ __forceinline__ bool ring_epilogue(uint32_t taddr, uint8_t *base_lane, uint64_t ptr_own, uint32_t bias_own, uint32_t bias_flag,
                                              uint32_t lane_off, uint32_t knockout, uint32_t keep_mask, uint64_t *release_bar) {
    if (EPI == 3) return ring_epilogue_w16o<FAST, STRIDE, BIAS, MASK>(taddr, base_lane, ptr_own, bias_own, bias_flag, lane_off, knockout, keep_mask, release_bar);
    if (EPI == 2) return ring_epilogue_burst<FAST, STRIDE, BIAS, MASK>(taddr, base_lane, ptr_own, bias_own, bias_flag, lane_off, knockout, keep_mask, release_bar);
    if (EPI == 1) return ring_epilogue_w16<FAST, STRIDE, BIAS, MASK>(taddr, base_lane, ptr_own, bias_own, bias_flag, lane_off, knockout, keep_mask, release_bar);
    if (knockout & 4u) return false;
    uint32_t v[2][4][8];
    ring_ld_chunk(taddr, v[0][0], v[0][1], v[0][2], v[0][3]);
#pragma unroll
    for (uint32_t g = 0; g < 4; ++g) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (g + 1 < 4) ring_ld_chunk(taddr + (g + 1) * 8u, v[(g + 1) & 1][0], v[(g + 1) & 1][1], v[(g + 1) & 1][2], v[(g + 1) & 1][3]);
        else if (release_bar) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) mbar_arrive(release_bar);
        }
#pragma unroll
        for (uint32_t c4 = 0; c4 < 8; c4 += 4) {
            const uint4 bq = ring_bias4<BIAS>(bias_own, g * 8u + c4);
#pragma unroll
            for (uint32_t c = c4; c < c4 + 4; ++c) {
                const uint32_t x = ((v[g & 1][3][c] * 256u + v[g & 1][2][c]) * 256u + v[g & 1][1][c]) * 256u + v[g & 1][0][c];
                ring_emit<FAST, STRIDE, BIAS, MASK>(g * 8u + c, x, base_lane, ptr_own, ring_q(bq, c - c4), bias_flag, lane_off, knockout, keep_mask);
            }
        }
    }
    return release_bar != nullptr;
}

__device__ __forceinline__ void stg128_zero_stream(void *p) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(0u) : "memory");
}

// ROT = NUM_REGIONS > 1 (rotated loads, masked b tail), BATCHED = several input / output sets in one launch: the plain
// instantiation carries none of that code (a run-time `batched` flag alone cost the single-set launch 5 %)
// EPI: epilogue schedule (0 x8 loads double-buffered against the stores, 1 x16 loads, 2 burst: all loads, release the stage, stores)
template <bool ROT, bool BATCHED, int EPI>
__global__ void __launch_bounds__(RG_THREADS, 1) cloud_ring_kernel(const RingParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // The production library compiles ONE schedule: the switches are constants there and every other path (knock-outs, tracing, the
    // alternative wait / issue variants) is dead code the compiler removes -- the kernel is ~6500 instructions with them, and five
    // roles running different parts of it share the SM's instruction cache.
#ifdef IDASH_B200_PROFILE
    const uint32_t k_tune = p.tune, k_knockout = p.knockout, k_trace = p.trace_cta;
#else
    // (NUM_REGIONS > 1 keeps the schedule word in a register: with it folded, ptxas allocates the rotated-staging producers -- five
    // 128-bit loads per feature, two blocks in flight -- differently and spills 76 bytes at the 128-register cap, which costs 60 %)
    const uint32_t k_tune = ROT ? p.tune : RG_TUNE_DEFAULT;
    constexpr uint32_t k_knockout = 0u, k_trace = 0u;
#endif
    __shared__ __align__(8) uint64_t a_full[RG_MAX_SLOTS], b_full[RG_BBARS], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t tiles_done_s;       // tiles of this CTA whose MMAs are complete
    __shared__ uint32_t blocks_freed_s;     // input blocks (in staging order) that no pending MMA reads any more
    __shared__ uint32_t mma_issued_s;       // tiles of this CTA whose MMAs have all been handed to the tensor pipe (tune & 32)
    __shared__ uint32_t waits_done_s;       // tiles of this CTA whose input-block waits have all been satisfied (see the MMA issuers)

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // CTAs [0, n_slices * n_chunks) are the (slice, chunk) pairs that fill the GPU. The SMs that division leaves over (148 = 16 x 9 + 4)
    // get one more, SHORT chunk at the end of the tile list (extra_tiles > 0: n_slices more CTAs): its CTAs run in waves on the spare SMs
    // -- the first few at once, the others as those retire -- so the chunk is 1 / waves as long as a regular one and all SMs finish
    // together. (Measured: the kernel's time follows the number of SMs it uses -- 9 / 8 / 7 chunks at neighbors = 5: 0.454 / 0.499 /
    // 0.544 ms, profiles/r02_ring_sm_count.txt.)
    const uint32_t chunk = blockIdx.x / p.n_slices, slice = blockIdx.x - chunk * p.n_slices;
    const bool extra = chunk >= p.n_chunks;
    // virtual tiles of this chunk: v in [t_begin, t_end), real tile = tile_base + v % n_tiles, batch = v / n_tiles
    const uint32_t n_virtual = p.n_tiles * p.n_batches, n_main = n_virtual - p.extra_tiles;
    const uint32_t t_begin = extra ? n_main : (uint32_t) ((uint64_t) n_main * chunk / p.n_chunks);
    const uint32_t t_end = extra ? n_virtual : (uint32_t) ((uint64_t) n_main * (chunk + 1) / p.n_chunks);
    if (t_begin >= t_end) return;
    constexpr bool batched = BATCHED;
    uint32_t *hdr_s = reinterpret_cast<uint32_t *>(smem + p.hdr_off);
    const uint32_t *const hdr_s_ = hdr_s;
    // (after the header table is built: the batch of a virtual tile of this chunk is read back from its header -- the roles
    // below map tiles to batches once per tile, and an integer division there cost the epilogue warps 18 % at neighbors = 5)
    auto batch_of = [&](uint32_t v) -> uint32_t { return batched ? (hdr_s_[v - t_begin] & ((1u << RG_HDR_A_BITS) - 1u)) >> RG_BATCH_SHIFT : 0u; };
    auto real_tile = [&](uint32_t v) -> uint32_t { return p.tile_base + v - batch_of(v) * p.n_tiles; };
    uint8_t *sA = smem;
    uint8_t *sB = smem + p.n_slots * RG_BLOCK_BYTES;

    const uint32_t n_my = t_end - t_begin;             // <= p.max_chunk_tiles
    for (uint32_t i = tid; i < n_my; i += RG_THREADS) {
        const uint32_t v = t_begin + i, b = batched ? v / p.n_tiles : 0u;
        hdr_s[i] = ring_pack_hdr((batched ? p.batch_tiles[b] : p.tiles) + p.tile_base + (v - b * p.n_tiles)) + (b << RG_BATCH_SHIFT);
    }
    if (ROT) {
        for (uint32_t i = tid; i < RG_ZERO_BYTES / 16u; i += RG_THREADS) reinterpret_cast<uint4 *>(smem + p.zero_off)[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // read by bulk stores (async proxy) after the barrier below
    }
    const uint32_t w_slice = slice * 128u;
    const bool is_b = (w_slice & POLY_N) != 0 && !(k_knockout & 32u);     // knock-out 32: every slice runs the (cheaper) epilogue of polynomial a
    const uint32_t i_slice = w_slice & (POLY_N - 1);

    if (tid == 0) {
        for (uint32_t s = 0; s < p.n_slots; ++s) mbar_init(&a_full[s], 4);
        for (uint32_t s = 0; s < RG_BBARS; ++s) mbar_init(&b_full[s], 1);
        for (uint32_t s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], RG_EPI_WARPS + 1); }   // 8 epilogue warps + the publisher
        tiles_done_s = 0;
        blocks_freed_s = 0;
        mma_issued_s = 0;
        waits_done_s = 0;
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
        const uint32_t keep_mask = (is_b && i_slice + word_in_slice >= p.RS) ? 0u : 0xFFFFFFFFu;
        const bool records = p.out.records != 0;
        const bool early = (k_tune & 64u) != 0u;
        // NUM_REGIONS > 1: the 16 - n_slices slices that lie entirely in b[RS..N) are not computed by anyone; the epilogue warps
        // of the n_slices computing CTAs zero-fill them, 512 bytes (one slice of one row) per warp instruction, tile by tile, so
        // that all 8 KB of an output ciphertext are still written at about the same time
        const uint32_t n_zero_seg = 16u - p.n_slices;
        const uint32_t zero_units = n_zero_seg * TC_TN, zero_workers = p.n_slices * RG_EPI_WARPS, zero_me = slice * RG_EPI_WARPS + warp;
        // Row information (caller rows + Constants of the tile) is in the tile's metadata record in shared memory, put there by the
        // coefficient loader's bulk copies together with the coefficient image (b_full of the tile covers both). The epilogue issues
        // no global load: one issued here queues behind the warp's own store backlog (measured: a load issued two tiles -- ~4000
        // cycles -- ahead still stalled its consumer for ~120 cycles per tile).
        const uint32_t bias_mul = bias_flag ? (uint32_t) IDASH_B200_ONE_IN_T32 : 0u;
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t it = t - t_begin, st = it & 1u;
            const uint32_t *meta = reinterpret_cast<const uint32_t *>(smem + p.meta_off + (it & (RG_META_STAGES - 1u)) * (8u * TC_TN));
            // the record landed before the tile's MMAs were issued; observing its barrier makes the bulk copy's bytes visible here
            mbar_spin(&b_full[it & (RG_BBARS - 1u)], (it / RG_BBARS) & 1u);
            const uint32_t row = meta[col_base + lane];
            const uint32_t bias_addr = smem_u32(meta + TC_TN + col_base);
            const bool fast = ring_hdr(hdr_s, it).fast && p.slot_of_row == nullptr;
            uint64_t ptr_own = 0;
            uint8_t *base_lane = nullptr;
            const uint32_t bt_ = batched ? batch_of(t) : 0u;
            uint8_t *const out_words = batched ? reinterpret_cast<uint8_t *>(p.batch_out[bt_]) : p.out.words;
            // (batched launches: the batch's own NUM_SAMPLES decides which words get the Constant)
            const uint32_t bias_mul_t = !batched ? bias_mul : ((is_b && i_slice + word_in_slice < p.batch_S[bt_]) ? (uint32_t) IDASH_B200_ONE_IN_T32 : 0u);
            if (fast) {
                const uint32_t row0 = __shfl_sync(0xFFFFFFFFu, row, 0);     // caller row of tile row col_base
                base_lane = out_words + (uint64_t) row0 * p.out.stride + 4u * w_slice + lane_off;
            } else if (row != IDASH_B200_NO_ROW) {
                ptr_own = (uint64_t) (out_words + (uint64_t) (p.slot_of_row ? __ldg(p.slot_of_row + row) : row) * p.out.stride + 4u * w_slice);
            }
            if (k_tune & 1u) mbar_wait(&t_full[st], (it >> 1) & 1u); else mbar_spin(&t_full[st], (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) RG_TRACE(6, it);
            const uint32_t taddr = tmem + ((quad * 32u) << 16) + st * 4u * TC_TN + col_base;
            uint64_t *const rel = early ? &t_empty[st] : nullptr;
            bool arrived;
            if (fast) {
                if (records) {
                    if (is_b) arrived = ring_epilogue<EPI, true, IDASH_B200_RECORD_BYTES, true, ROT>(taddr, base_lane, 0, bias_addr, bias_mul_t, lane_off, k_knockout, keep_mask, rel);
                    else arrived = ring_epilogue<EPI, true, IDASH_B200_RECORD_BYTES, false>(taddr, base_lane, 0, 0, 0, lane_off, k_knockout, 0xFFFFFFFFu, rel);
                } else {
                    if (is_b) arrived = ring_epilogue<EPI, true, IDASH_B200_CT_BYTES, true, ROT>(taddr, base_lane, 0, bias_addr, bias_mul_t, lane_off, k_knockout, keep_mask, rel);
                    else arrived = ring_epilogue<EPI, true, IDASH_B200_CT_BYTES, false>(taddr, base_lane, 0, 0, 0, lane_off, k_knockout, 0xFFFFFFFFu, rel);
                }
            } else {
                arrived = ring_epilogue<EPI, false, 0, true, ROT>(taddr, nullptr, ptr_own, bias_addr, bias_mul_t, lane_off, k_knockout, keep_mask, rel);
            }
            if (!arrived) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[st]);      // one arrival per epilogue warp
            }
            if (ROT && zero_units && !(k_tune & 16u) && !(k_knockout & (2u | 16u))) {
                // Zero fill of b[128 n_slices - 1024 .. 1024) of this tile's rows: the segments of a row are contiguous (2560 bytes at
                // REGION_SIZE = 341), so a row is ONE bulk store from the block of zeros in shared memory, issued by one lane and carried
                // out by the copy engine -- as 128-bit stores of all lanes the fill took 15 % of the kernel (knock-out 16,
                // profiles/r02_ring_knockout.txt). The tile's rows go round the n_slices x 8 epilogue warps of the chunk.
                const uint32_t row_lane0 = __shfl_sync(0xFFFFFFFFu, row, 0);       // caller row of tile row col_base
                if (lane == 0) {
                    const uint32_t first = (zero_me + zero_workers - it % zero_workers) % zero_workers;
                    for (uint32_t n = first; n < TC_TN; n += zero_workers) {
                        uint32_t r = fast ? row_lane0 - col_base + n : meta[n];
                        if (r == IDASH_B200_NO_ROW) continue;
                        if (p.slot_of_row) r = __ldg(p.slot_of_row + r);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(out_words + (uint64_t) r * p.out.stride + 512u * p.n_slices), "r"(smem_u32(smem + p.zero_off)), "r"(512u * n_zero_seg) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (ROT && zero_units && !(k_knockout & (2u | 16u))) {
                // (tune & 16: the same fill as 512-byte units of 128-bit stores, unit u = (tile row u / n_zero_seg, segment u % n_zero_seg))
                const uint32_t row_lane0 = __shfl_sync(0xFFFFFFFFu, row, 0);       // caller row of tile row col_base
                for (uint32_t u = zero_me; u < zero_units; u += zero_workers) {
                    const uint32_t n = u / n_zero_seg, seg = u - n * n_zero_seg;
                    uint32_t r = fast ? row_lane0 - col_base + n : meta[n];
                    if (r == IDASH_B200_NO_ROW) continue;
                    if (p.slot_of_row) r = __ldg(p.slot_of_row + r);
                    stg128_zero_stream(out_words + (uint64_t) r * p.out.stride + 512u * (p.n_slices + seg) + 16u * lane);
                }
            }
            if (tid == 0) RG_TRACE(8, it);
        }
        if (ROT && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // the zero-fill stores of this lane are complete
    } else if (warp == RG_WARP_MMA || warp == RG_WARP_MMA2) {
        // ================= MMA issuers =================
        // Two warps: warp RG_WARP_MMA issues the even tiles of the chunk (TMEM stage 0), RG_WARP_MMA2 the odd ones (stage 1).
        // One issuing thread is a serial instruction stream of a few hundred instructions per tile (waits, operand set-up,
        // MMAs, commit, ring bookkeeping) and was what bounded the tile rate; the two stages are independent chains, so
        // they get one stream each. Both warps walk every tile to keep the ring state, but wait and issue only for their own.
        const uint32_t q = warp == RG_WARP_MMA2 ? 1u : 0u;
        const uint32_t n_slots = p.n_slots, n_bchunks = p.n_bchunks;
        const bool fifo = (k_tune & 32u) != 0u;
        const uint32_t stagger = ((k_tune >> 12) & 15u) ? min((k_tune >> 12) & 15u, 8u) : 8u;     // eighths of a tile, see pub_ks below
        const uint64_t da_base = tc_desc(smem_u32(sA), TC_A_LBO, TC_A_SBO);
        const uint64_t db_base = tc_desc(smem_u32(sB), TC_B_LBO, TC_B_SBO);
        uint32_t it = 0;
        uint32_t next_slot = 0, next_par = 0;       // slot / phase parity of the next input block to be staged
        uint32_t first_slot = 0;                    // slot of block T.a
        uint32_t bpos = 0;                          // coefficient ring position (chunk) of the tile's first K step
        uint32_t prev_fn = 0, prev_bt = 0, prev_slot = 0, prev_par = 0;   // new blocks of the previous tile: [prev_fn, prev_bt) from prev_slot
        RingWalk walk;
        walk.init(ring_hdr(hdr_s, 0).a);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        const uint32_t leader = elect_one();
        for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
            const bool has_next = t + 1 < t_end;
            const RingTile T = ring_hdr(hdr_s, it), Tn = ring_hdr(hdr_s, has_next ? it + 1u : it);
            const uint32_t st = it & 1u;
            const bool mine = st == q;
            const uint32_t bt = T.a + T.nb;
            const uint32_t fn = walk.first_new(T);
            const uint32_t n_new = bt - fn;
            if (mine) {
                if (lane == 0) RG_TRACE(0, it);
                // all waits of this tile at once, one barrier per lane: lane 0 the TMEM stage, lane 1 the coefficient stage,
                // then the input blocks this tile adds to the ring, then those the previous tile (issued by the other warp)
                // added and this tile uses
                const uint32_t p_lo = max(T.a, prev_fn), p_hi = min(prev_bt, bt);
                const uint32_t n_prev = p_hi > p_lo ? p_hi - p_lo : 0u;
                uint64_t *bar = nullptr;
                uint32_t par = 0;
                if (lane == 0) { bar = &t_empty[st]; par = ((it >> 1) & 1u) ^ 1u; }
                else if (lane == 1) { bar = &b_full[it & (RG_BBARS - 1u)]; par = (it / RG_BBARS) & 1u; }
                else if (lane - 2u < n_new) {
                    uint32_t s = next_slot + (lane - 2u);
                    par = next_par;
                    if (s >= n_slots) { s -= n_slots; par ^= 1u; }
                    bar = &a_full[s];
                } else if (lane - 2u - n_new < n_prev) {
                    uint32_t s = prev_slot + (p_lo - prev_fn) + (lane - 2u - n_new);
                    par = prev_par;
                    if (s >= n_slots) { s -= n_slots; par ^= 1u; }
                    bar = &a_full[s];
                }
                // The a_full waits are PARITY waits on ring slots that are reused every n_slots blocks, and a parity wait is only
                // meaningful when the slot's previous use is known to be complete: test_wait(parity of use k) is also true while
                // the barrier is still in use k - 1. One warp waiting for its tiles in order would guarantee that; two warps on
                // alternating tiles do not -- this warp may get here before the blocks of the OTHER warp's previous tile (whose slots
                // this tile's new blocks may reuse when the band moves by several blocks per tile) have been staged at all, and in
                // particular at kernel start, when nothing has (found by running the 9-set NUM_REGIONS = 2 test under
                // compute-sanitizer: with the producers slowed down, the last rows of a chunk were computed from a stale slot).
                // Gate: the input-block lanes start polling only when the previous tile's input waits are all satisfied (waits_done_s);
                // then every block staged before this tile's is in place, i.e. the previous use of every slot this tile adds to.
                // It is needed only when a slot's previous occupant (n_slots blocks earlier in staging order) is not below what this
                // warp has itself seen staged, i.e. when the new blocks of this tile and the previous one together wrap the ring --
                // never on the iDASH geometry (less than one new block per tile), and then neither the gate nor the publication
                // costs an instruction in the polling loop (gating / publishing every tile cost 1.7 % at neighbors = 5 and 4.3 %
                // at 50, A/B on one box).
                // waits_done_s is published as soon as the INPUT lanes are through -- the other warp's input polling need not wait for this
                // tile's TMEM stage or coefficients.
                // gate(it) <=> the new blocks of tiles it - 1 and it together wrap the ring; then (and only then) tile it - 1 publishes
                const uint32_t st_after = max(walk.staged_upto, bt);
                const uint32_t n_new_next = has_next ? Tn.a + Tn.nb - max(Tn.a, st_after) : 0u;
                const bool need_gate = (prev_bt - prev_fn) + n_new > n_slots && !(k_knockout & 128u), need_pub = n_new + n_new_next > n_slots;     // knock-out 128: no gate
                bool gate_open = !need_gate, published = !need_pub;
                // warp-uniform polling loop: lanes without a barrier count as done
                const uint32_t bar_addr = bar ? smem_u32(bar) : 0u;
                // lane 31 (tune & 32): the other warp has handed ALL MMAs of the previous tile to the tensor pipe. The pipe executes
                // in issue order; without this the two warps' MMAs interleave, both tiles complete late and at the same time, and
                // the two TMEM stages run in step (MMA phase, then epilogue phase) instead of overlapping.
                const bool order_lane = fifo && lane == 31u && it != 0u;
                uint32_t done = (bar || order_lane) ? 0u : 1u;
                long long t_done = 0;
                if (k_tune & 1024u) {
                    // every lane blocks on its own barrier with try_wait (the hardware suspends the thread until the phase completes or a
                    // time limit expires): no polling granularity between "TMEM stage released" and the first MMA of the next tile
                    if (need_gate) progress_wait(&waits_done_s, it);
                    if (lane >= 2u && bar) mbar_wait(bar, par);
                    __syncwarp();
                    if (need_pub && lane == 0) progress_publish_max(&waits_done_s, it + 1u);
                    published = true;
                    if (lane < 2u && bar) mbar_wait(bar, par);
                    if (order_lane) {
                        uint32_t v;
                        do asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&mma_issued_s)) : "memory");
                        while ((int32_t) (v - it) < 0);
                    }
                    if (k_trace) t_done = clock64();
                    __syncwarp();
                } else
                for (;;) {
                    if (!gate_open) {
                        uint32_t v;
                        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&waits_done_s)) : "memory");
                        gate_open = (int32_t) (v - it) >= 0;
                    }
                    if (!done) {
                        if (order_lane) {
                            uint32_t v;
                            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&mma_issued_s)) : "memory");
                            done = (int32_t) (v - it) >= 0 ? 1u : 0u;
                        } else if (lane < 2u || gate_open) {
                            done = mbar_test(bar_addr, par);
                        }
                        if (done && k_trace) t_done = clock64();
                    }
                    // (published with a max: a tile that needed no gate may get here before the other warp's previous tile, and a plain
                    // store of that tile's `it` would then overwrite `it + 2` -- the first version of this gate hung on exactly that)
                    if (!published && gate_open && __all_sync(0xFFFFFFFFu, done || lane < 2u)) {
                        if (lane == 0) progress_publish_max(&waits_done_s, it + 1u);      // opens the other warp's gate for tile it + 1
                        published = true;
                    }
                    if (published && gate_open && __all_sync(0xFFFFFFFFu, done)) break;
                    if (!(k_tune & 4u)) __nanosleep(40);      // polling hot next to running MMAs slows them down (measured)
                }
                if (k_trace) {
                    if (lane == 0) { RG_TRACE_V(1, it, t_done); RG_TRACE(4, it); }
                    if (lane == 1) RG_TRACE_V(2, it, t_done);
                    if (lane == 2) RG_TRACE_V(3, it, t_done);
                }
            }
            prev_fn = fn; prev_bt = bt; prev_slot = next_slot; prev_par = next_par;
            next_slot += n_new;
            if (next_slot >= n_slots) { next_slot -= n_slots; next_par ^= 1u; }
            if (mine) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (k_tune & 8u) {
                // warp-converged issue: every operand is made provably warp-uniform (shuffle from lane 0), so that the
                // compiler feeds the MMA's uniform-register operands without a per-instruction broadcast loop
                const uint32_t nb_u = __shfl_sync(0xFFFFFFFFu, T.nb, 0);
                uint32_t aslot = __shfl_sync(0xFFFFFFFFu, first_slot, 0);
                const uint32_t d0 = tmem_u + st * 4u * TC_TN;
                uint32_t bchunk = __shfl_sync(0xFFFFFFFFu, bpos, 0);
                // staggered hand-over (tune & 32): the other warp may start on tile it + 1 once stagger / 8 of this tile's K steps have been
                // handed to the tensor pipe (8 / 8 = strictly one tile after the other)
                const uint32_t pub_ks = fifo ? (nb_u * stagger + 7u) / 8u - 1u : 0xFFFFFFFFu;
                if (lane == 0) RG_TRACE(7, it);
                if (!(k_knockout & 1u)) {
                    if (k_tune & 128u) {
                        for (uint32_t ks = 0; ks < nb_u; ++ks) {
                            const uint64_t da = da_base + (uint64_t) ((aslot * RG_BLOCK_BYTES) >> 4), db = db_base + (uint64_t) ((bchunk * TC_B_CHUNK) >> 4);
                            if (q) tc_mma_kstep_ws_p<1>(d0, da, RG_PLANE_BYTES >> 4, db, ks == 0, leader);
                            else tc_mma_kstep_ws_p<0>(d0, da, RG_PLANE_BYTES >> 4, db, ks == 0, leader);
                            if (ks == pub_ks && leader) progress_publish(&mma_issued_s, it + 1u);
                            if (++aslot == n_slots) aslot = 0;
                            if (++bchunk == n_bchunks) bchunk = 0;
                        }
                    } else
                    for (uint32_t ks = 0; ks < nb_u; ++ks) {
                        tc_mma_kstep_p(d0, da_base + (uint64_t) ((aslot * RG_BLOCK_BYTES) >> 4), RG_PLANE_BYTES >> 4,
                                       db_base + (uint64_t) ((bchunk * TC_B_CHUNK) >> 4), ks == 0, leader);
                        if (ks == pub_ks && leader) progress_publish(&mma_issued_s, it + 1u);
                        if (++aslot == n_slots) aslot = 0;
                        if (++bchunk == n_bchunks) bchunk = 0;
                    }
                }
                if (lane == 0) RG_TRACE(11, it);
                if (leader) {
                    if (fifo) progress_publish(&mma_issued_s, it + 1u);
                    tc_commit(&t_full[st]);
                }
                if (lane == 0) RG_TRACE(5, it);
            } else
            if (lane == 0) {
                const uint32_t d0 = tmem + st * 4u * TC_TN;
                uint32_t aslot = first_slot, bchunk = bpos;
                if (!(k_knockout & 1u)) {
                    for (uint32_t ks = 0; ks < T.nb; ++ks) {
                        tc_mma_kstep(d0, da_base + (uint64_t) ((aslot * RG_BLOCK_BYTES) >> 4), RG_PLANE_BYTES >> 4,
                                     db_base + (uint64_t) ((bchunk * TC_B_CHUNK) >> 4), ks == 0);
                        if (++aslot == n_slots) aslot = 0;
                        if (++bchunk == n_bchunks) bchunk = 0;
                    }
                }
                if (fifo) progress_publish(&mma_issued_s, it + 1u);
                tc_commit(&t_full[st]);     // the only commit of the tile: epilogue and publisher wait on it
                RG_TRACE(5, it);
            }
            __syncwarp();
            }
            bpos += T.nb;
            if (bpos >= n_bchunks) bpos -= n_bchunks;
            uint32_t rb, re;
            walk.advance(T, Tn, has_next, rb, re);
            // slot of the next tile's first block
            if (has_next) {
                if (Tn.a >= walk.staged_upto) first_slot = next_slot;              // gap: its blocks are all new
                else { first_slot += Tn.a - T.a; while (first_slot >= n_slots) first_slot -= n_slots; }
            }
        }
    } else if (warp == RG_WARP_BLOAD) {
        // ================= coefficient loader: the tile's image into a ring of 4096-byte chunks =================
        // A tile takes K / 32 consecutive chunks (wrapping), so narrow tiles do not reserve the widest tile's size and the image of
        // tile t + 2 can be on its way while tiles t and t + 1 still hold theirs. One mbarrier per tile (b_full[it % 8]) collects the
        // bytes of its (at most two) bulk copies.
        if (lane == 0) {
            uint32_t it = 0, cpos = 0, loaded = 0, freed = 0, done_it = 0;
            // the coefficient images of consecutive tiles are contiguous (layout.cpp): b_off advances by the tile's size
            uint32_t lb = batched ? batch_of(t_begin) : 0u;            // batch (= model) of the tile being loaded
            const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>((batched ? p.batch_tiles[lb] : p.tiles) + real_tile(t_begin)));
            uint64_t b_off = (uint64_t) h0.z | ((uint64_t) h0.w << 32);
            for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
                const RingTile T = ring_hdr(hdr_s, it);
                if (batched && t != t_begin && real_tile(t) == p.tile_base) {     // next batch: back to the first image (of that batch's model)
                    lb = batch_of(t);
                    const uint4 hb = __ldg(reinterpret_cast<const uint4 *>(p.batch_tiles[lb] + p.tile_base));
                    b_off = (uint64_t) hb.z | ((uint64_t) hb.w << 32);
                }
                const uint8_t *const coef_base = batched ? p.batch_coef[lb] : p.tile_coef;
                // Barrier and metadata record it % 8 were last used by tile it - 8, whose epilogue reads the record until its last
                // store. The MMAs of tile it - 4 being complete implies that: they needed the TMEM stage of tile it - 6, which the
                // epilogue warps hand back only after they have finished tiles it - 7 and it - 8.
                if (it >= 4u) progress_wait(&tiles_done_s, it - 3u);
                const uint32_t bytes = T.nb * TC_B_CHUNK;
                uint64_t *const bar = &b_full[it & (RG_BBARS - 1u)];
                RG_TRACE(10, it);
                mbar_arrive_expect_tx(bar, bytes + 8u * TC_TN);
                {   // metadata record: caller rows, then Constants, of the tile's 64 rows (256 bytes each, consecutive tiles contiguous)
                    const uint32_t tt = real_tile(t);
                    uint8_t *const rec = smem + p.meta_off + (it & (RG_META_STAGES - 1u)) * (8u * TC_TN);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(rec)), "l"((batched ? p.batch_rows[lb] : p.tile_rows) + (uint64_t) tt * TC_TN), "r"(4u * TC_TN), "r"(smem_u32(bar)) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(rec + 4u * TC_TN)), "l"((batched ? p.batch_bias[lb] : p.tile_bias) + (uint64_t) tt * TC_TN), "r"(4u * TC_TN), "r"(smem_u32(bar)) : "memory");
                }
                // Ring space is admitted chunk by chunk: whatever part of the image fits is copied NOW and the rest as soon as earlier
                // tiles complete. (Waiting for room for the whole image put its copy -- ~1800 cycles for 28 KB at neighbors = 50, where
                // three images are one chunk more than the ring holds -- on the critical path of every tile.)
                uint32_t done_chunks = 0;
                while (done_chunks < T.nb) {
                    uint32_t room = p.n_bchunks - (loaded - freed);
                    while (room == 0u) {
                        progress_wait(&tiles_done_s, done_it + 1u);
                        freed += ring_hdr(hdr_s, done_it).nb;
                        ++done_it;
                        room = p.n_bchunks - (loaded - freed);
                    }
                    // free chunks that are already known to be free (no waiting): take as much as possible in one copy
                    const uint32_t n_now = min(min(T.nb - done_chunks, room), p.n_bchunks - cpos);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(sB + cpos * TC_B_CHUNK)), "l"(coef_base + b_off + (uint64_t) done_chunks * TC_B_CHUNK),
                                   "r"(n_now * TC_B_CHUNK), "r"(smem_u32(bar))
                                 : "memory");
                    done_chunks += n_now;
                    loaded += n_now;
                    cpos += n_now;
                    if (cpos >= p.n_bchunks) cpos -= p.n_bchunks;
                }
                b_off += bytes;
                // The images are contiguous, so the ones a few tiles ahead are pulled into L2 now: a copy that starts when ring space
                // frees up then pays the L2 latency, not HBM's (measured at neighbors = 50: 0.673 -> 0.640 ms).
                if (p.coef_prefetch) {
                    const uint64_t pf = b_off + (uint64_t) p.coef_prefetch * bytes;
                    if (pf + bytes <= (batched ? p.batch_coef_bytes[lb] : p.coef_bytes))
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(coef_base + pf), "r"(bytes) : "memory");
                }
            }
        }
    } else if (warp == RG_WARP_PUB) {
        // ================= progress publisher =================
        // Sees each tile's t_full complete (= its MMAs are done) and advances the two counters the loader and
        // the producers poll, so that the MMA thread needs no per-slot / per-stage tcgen05.commit.
        if (lane == 0) {
            RingWalk walk;
            walk.init(ring_hdr(hdr_s, 0).a);
            uint32_t freed = 0, it = 0;
            for (uint32_t t = t_begin; t < t_end; ++t, ++it) {
                const RingTile T = ring_hdr(hdr_s, it), Tn = ring_hdr(hdr_s, t + 1 < t_end ? it + 1u : it);
                uint32_t rb, re;
                walk.advance(T, Tn, t + 1 < t_end, rb, re);
                freed += re - rb;
                if (k_tune & 2u) mbar_wait(&t_full[it & 1u], (it >> 1) & 1u); else mbar_spin(&t_full[it & 1u], (it >> 1) & 1u);
                RG_TRACE(9, it);
                progress_publish(&blocks_freed_s, freed);
                progress_publish(&tiles_done_s, it + 1u);
                mbar_arrive(&t_empty[it & 1u]);     // the stage's next phase cannot complete before this one was seen here
            }
        }
    } else {
        // ================= block producers (4 warps, 128 threads) =================
        const uint32_t ptid = tid - RG_WARP_PROD * 32u;
        const uint32_t mg = ptid & 7u, k0 = ptid >> 3;          // this thread stages features k0 and k0 + 16 of a block
        uint32_t slot = 0, seq = 0;
        // cursor over the input blocks of the chunk in staging order: tile by tile, the blocks [max(T.a, staged), T.a + T.nb)
        uint32_t cur_it = 0, cur_kb = 0, cur_bt = 0, staged_upto = ring_hdr(hdr_s, 0).a;
        auto next_block = [&]() -> uint32_t {
            while (cur_kb >= cur_bt) {
                if (cur_it >= n_my) return 0xFFFFFFFFu;
                const RingTile T = ring_hdr(hdr_s, cur_it++);
                cur_kb = max(T.a, staged_upto);
                cur_bt = T.a + T.nb;
                staged_upto = max(staged_upto, cur_bt);
            }
            return cur_kb++;
        };
        // global loads of one block: features k0 and k0 + 16, 16 words each. NUM_REGIONS > 1: feature f is ciphertext
        // f / NR multiplied by X^(-(f % NR) RS) (torusPolynomialMulByXai, toruspolynomial-functions.cpp:140-160): the words
        // are read rotated, with the sign flipped where the index wraps.
        // Rotated reads: words start .. start + 15 of the negacyclic extension, start = i0 + shift < 2047, as FIVE ALIGNED
        // 128-bit loads (REGION_SIZE = 341 makes two of three features unaligned; sixteen 32-bit loads per feature made the
        // producers' load/store unit the bottleneck). load_block only ISSUES the loads and keeps the raw groups; the sign
        // flip and the word-granular realignment by start % 4 happen in store_block -- consuming the data right after
        // the loads made every block wait out its own HBM latency and arrive ~2000 cycles after its tile was ready.
        constexpr int NW = ROT ? 5 : 4;
        auto rot_start = [&](uint32_t f) -> uint32_t {       // first word of this thread's window in the negacyclic extension
            const uint32_t ct = f / p.NR;
            return i_slice + mg * 16u + (f - ct * p.NR) * p.RS;
        };
        auto load_block = [&](uint32_t kbv, uint4 (&w)[2][NW]) {
            const uint32_t kb = batched ? kbv & ((1u << RG_BATCH_SHIFT) - 1u) : kbv;
            const uint8_t *const in_words = batched ? reinterpret_cast<const uint8_t *>(p.batch_in[kbv >> RG_BATCH_SHIFT]) : p.in.words;
            const uint32_t pb = batched ? kbv >> RG_BATCH_SHIFT : 0u;
            const uint32_t used_word = kb < (batched ? p.batch_nfw[pb] : p.n_feat_words) ? __ldg((batched ? p.batch_feat_used[pb] : p.feat_used) + kb) : 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t k = k0 + 16u * h;
                const uint32_t f = kb * 32u + k;
                const uint32_t ct = ROT ? f / p.NR : f;
                uint32_t sl = NO_SLOT;
                if (ct < p.n_ct_slots) sl = p.slot_of_ct ? __ldg(p.slot_of_ct + ct) : ct;
                if (k_knockout & 8u) sl = NO_SLOT;
                if (sl == NO_SLOT) {
                    if (sl == NO_SLOT && ((used_word >> k) & 1u) && !(k_knockout & 8u)) atomicOr(p.status, 1);
#pragma unroll
                    for (int q = 0; q < NW; ++q) w[h][q] = make_uint4(0, 0, 0, 0);
                } else if (!ROT) {
                    const uint8_t *src = in_words + (uint64_t) sl * p.in.stride + 4u * (w_slice + mg * 16u);
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[h][q] = ldg128(src + 16 * q);
                } else {
                    const uint8_t *poly = in_words + (uint64_t) sl * p.in.stride + 4u * (w_slice & POLY_N);
                    const uint32_t start = rot_start(f), al = start & ~3u;
#pragma unroll
                    for (int q = 0; q < NW; ++q)
                        w[h][q] = (q < 4 || (start & 3u) != 0u) ? ldg128(poly + 4u * ((al + 4u * q) & (POLY_N - 1u))) : make_uint4(0, 0, 0, 0);
                }
            }
        };
        // split into byte planes and store into the ring slot, once the slot's previous block has been released
        auto store_block = [&](uint32_t kbv, const uint4 (&w)[2][NW]) {
            const uint32_t kb = batched ? kbv & ((1u << RG_BATCH_SHIFT) - 1u) : kbv;
            if (k_knockout & 64u) __nanosleep(20000);      // profiling build: slow producers (the timing compute-sanitizer produces; results stay right)
            if (seq >= p.n_slots) progress_wait(&blocks_freed_s, seq - p.n_slots + 1u);
            uint8_t *blk = sA + slot * RG_BLOCK_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t k = k0 + 16u * h;
                uint4 v[4];
                if (!ROT) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = w[h][q];
                } else {
                    const uint32_t start = rot_start(kb * 32u + k), o = start & 3u, al = start & ~3u;
                    uint32_t fl[20];
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const bool neg = ((al + 4u * j) & POLY_N) != 0u;     // an aligned group never straddles the wrap at 1024
                        uint4 x = w[h][j];
                        if (neg) { x.x = 0u - x.x; x.y = 0u - x.y; x.z = 0u - x.z; x.w = 0u - x.w; }
                        fl[4 * j] = x.x; fl[4 * j + 1] = x.y; fl[4 * j + 2] = x.z; fl[4 * j + 3] = x.w;
                    }
                    uint32_t t1[18];
#pragma unroll
                    for (int m = 0; m < 18; ++m) t1[m] = (o & 1u) ? fl[m + 1] : fl[m];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = make_uint4((o & 2u) ? t1[4 * q + 2] : t1[4 * q], (o & 2u) ? t1[4 * q + 3] : t1[4 * q + 1],
                                          (o & 2u) ? t1[4 * q + 4] : t1[4 * q + 2], (o & 2u) ? t1[4 * q + 5] : t1[4 * q + 3]);
                }
                uint32_t limb[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t x0 = __byte_perm(v[q].x, v[q].y, 0x5140), x1 = __byte_perm(v[q].x, v[q].y, 0x7362);
                    const uint32_t x2 = __byte_perm(v[q].z, v[q].w, 0x5140), x3 = __byte_perm(v[q].z, v[q].w, 0x7362);
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
            if (++slot == p.n_slots) slot = 0;
            ++seq;
        };
        // Two blocks in flight per thread (register sets A and B, statically alternated): the loads of the next block are
        // issued before the current one is split and stored, so the ~1 us HBM latency of a block overlaps the previous one.
        uint4 wA[2][NW], wB[2][NW];
        uint32_t kbA = next_block();
        if (kbA != 0xFFFFFFFFu) load_block(kbA, wA);
        while (kbA != 0xFFFFFFFFu) {
            const uint32_t kbB = next_block();
            if (kbB != 0xFFFFFFFFu) load_block(kbB, wB);
            store_block(kbA, wA);
            if (kbB == 0xFFFFFFFFu) break;
            kbA = next_block();
            if (kbA != 0xFFFFFFFFu) load_block(kbA, wA);
            store_block(kbB, wB);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == RG_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
