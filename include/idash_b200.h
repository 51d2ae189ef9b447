/* idash_b200.h -- C ABI of libidash_b200.so: the B200 (sm_100a) evaluator for the encrypted
 * genotype-imputation path of ssmiler/idash2019_2.
 *
 * The reference has no plugin/FFI layer; its boundary for this path is three C++ functions of
 * eval/idash.h compiled into the cloud / decrypt binaries:
 *     void cloud_compute_score(EncryptedPredictions&, const EncryptedData&, const Model&,
 *                              const IdashParams&);                     eval/idash.h:251-252
 *     void decrypt_predictions(DecryptedPredictions&, const EncryptedPredictions&,
 *                              const IdashKey&);                        eval/idash.h:254
 *     void read_model(Model&, const IdashParams&, const std::string&);  eval/idash.h:229
 * The entry points below are what new bodies of those functions bind (see INTEGRATION.md for the
 * exact glue): plain pointers and sizes, no C++ or torch types, int return codes, never abort().
 *
 * Conventions
 *   - one idash_b200_ctx per GPU (one process per GPU, or one ctx per GPU of a process that shards an evaluation). A ctx is not
 *     thread-safe and is used FROM ONE STREAM AT A TIME: the *_device entry points keep per-ctx scratch (ciphertext -> slot table,
 *     variance flags, status word, one fork / join event pair) that is not ordered against work the caller has queued on another
 *     stream. Two evaluations in flight on different streams need two contexts.
 *   - a TRLWE ciphertext ("ct") is 2048 uint32 words: polynomial a[1024] then b[1024]
 *     (tfhe/src/libtfhe/tlwe.cpp:42-52, k = 1, N = 1024, eval/idash.h:49-50).
 *   - all torus arithmetic is mod 2^32; results are bit-identical to the reference.
 *   - *_host entry points take HOST pointers (pinned recommended) and do the H2D/D2H copies;
 *     *_device entry points take DEVICE pointers, enqueue on the given cudaStream_t and return
 *     without synchronising.
 */
#ifndef IDASH_B200_H
#define IDASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDASH_B200_N 1024u               /* IdashParams::N, eval/idash.h:49 */
#define IDASH_B200_CT_WORDS 2048u        /* (k+1)*N */
#define IDASH_B200_CT_BYTES 8192u
#define IDASH_B200_RECORD_BYTES 8208u    /* u32 index, i32 84, 2048 words, f64 variance: eval/idash.cpp:551-552,
                                            tfhe/src/libtfhe/tfhe_io.cpp:315-323 */
#define IDASH_B200_TLWE_SAMPLE_UID 84    /* tfhe/src/include/tfhe_generic_streams.h:16 */
#define IDASH_B200_CONSTANT_BIDX 0xFFFFFFFFu /* IdashParams::constant_bigIndex(), eval/idash.h:93 */
#define IDASH_B200_ONE_IN_T32 262144     /* dtot32(1/16384), eval/idash.cpp:29-39 */

enum {
    IDASH_B200_OK = 0,
    IDASH_B200_ERR_INVALID = -1,       /* bad argument / inconsistent model */
    IDASH_B200_ERR_CUDA = -2,          /* CUDA runtime error (no device, launch failure, ...) */
    IDASH_B200_ERR_MISSING_INPUT = -3, /* model references a ciphertext that was not supplied
                                          (the reference aborts: eval/idash.h:164) */
    IDASH_B200_ERR_NOMEM = -4
};

enum {
    /* data = uint32[count][2048]; index / variance are separate arrays */
    IDASH_B200_LAYOUT_PACKED = 0,
    /* data = the record stream of encrypted_data.bin / encrypted_prediction.bin, i.e. file image + 8:
       count records of 8208 bytes {u32 index, i32 84, u32 a[1024], u32 b[1024], f64 variance}
       (eval/idash.cpp:540-556, 596-613). `data` itself is 8 mod 16, so the word arrays are 16-byte
       aligned; index / variance live inside the records and the two pointers below must be NULL. */
    IDASH_B200_LAYOUT_RECORDS = 1
};

/* A view of `count` ciphertexts (host or device memory, depending on the entry point). */
typedef struct idash_b200_cts {
    int32_t layout;
    void *data;
    uint64_t count;
    /* PACKED only. Inputs: ciphertext index (EncryptedData key = bigIndex / NUM_REGIONS,
       eval/idash.h:87,137) of every slot, NULL = slot i holds index i. Outputs: receives the output
       bigIndex of every slot, may be NULL. */
    uint32_t *index;
    /* PACKED only. TLweSample::current_variance of every slot. Inputs: NULL = alpha^2 = 2^-50 (what
       encrypt writes, eval/idash.cpp:20,625). Outputs: may be NULL. */
    double *variance;
} idash_b200_cts;

/* The Model of eval/idash.h:129-134 flattened to CSR; rows in any order, out_bidx unique. */
typedef struct idash_b200_model_desc {
    uint32_t num_samples;   /* IdashParams::NUM_SAMPLES  */
    uint32_t num_regions;   /* IdashParams::NUM_REGIONS  */
    uint32_t region_size;   /* IdashParams::REGION_SIZE  */
    uint64_t n_rows;        /* output features            */
    const uint32_t *out_bidx; /* [n_rows] output bigIndex */
    const uint64_t *row_ptr;  /* [n_rows+1]               */
    const uint32_t *col;      /* [nnz] input bigIndex or IDASH_B200_CONSTANT_BIDX */
    const int32_t *coef;      /* [nnz]                    */
} idash_b200_model_desc;

typedef struct idash_b200_model_info {
    uint64_t n_rows, nnz;
    uint64_t n_groups;        /* device row groups (two target SNPs x three variants each) */
    uint64_t n_entries;       /* distinct (ciphertext, rotation) pairs summed over groups */
    uint32_t ct_min, ct_max;  /* range of input ciphertext indices the model touches (ct_min > ct_max: none) */
    uint32_t max_entries_per_group;
    uint32_t shifts_aligned;  /* 1 if every rotation is a multiple of 4 words (128-bit path) */
    uint64_t device_bytes;    /* size of the device-resident layout */
    uint64_t n_tiles;         /* band tiles of the tensor-core kernel; 0 = model not eligible (see idash_b200_layout.h) */
    uint32_t tile_kmax;       /* widest tile band (features) */
    uint32_t ring_ok;         /* 1 if the persistent ring variant of the tensor-core kernel applies */
    uint64_t n_overflow_rows; /* rows no tile holds (evaluated by the IMAD kernel beside the tensor-core kernel) */
    uint32_t groups_all;      /* 1 if the IMAD groups of every row have been built (n_groups / n_entries then count those) */
    uint32_t pad;
} idash_b200_model_info;

typedef struct idash_b200_ctx idash_b200_ctx;
typedef struct idash_b200_model idash_b200_model;

/* ---- context --------------------------------------------------------------------------------- */
int idash_b200_init(idash_b200_ctx **ctx, int device);
int idash_b200_destroy(idash_b200_ctx *ctx);
/* text of the last error on the calling thread ("" if none) */
const char *idash_b200_last_error(void);
/* number of kernels this ctx has launched so far */
uint64_t idash_b200_kernel_launches(const idash_b200_ctx *ctx);
/* Per-launch device timing of the dominant kernel (cloud_eval_kernel / decrypt_kernel): when enabled,
 * every launch is bracketed by cudaEvents on its own stream (up to max_launches, then it stops
 * recording). timing_read() synchronises those events and returns the durations in milliseconds in
 * launch order; *n is in: capacity of ms[], out: launches recorded. timing_enable(ctx, 0) disables. */
int idash_b200_timing_enable(idash_b200_ctx *ctx, int max_launches);
int idash_b200_timing_read(idash_b200_ctx *ctx, float *ms, int *n);
/* Which cloud kernel idash_b200_cloud_eval_* launches. AUTO = a tensor-core kernel (tcgen05 int8 limb-split
 * GEMM over band tiles) when the model is eligible -- every coefficient fits int16 and every 64-row tile's band
 * is at most 256 features wide -- else the IMAD kernel. TENSOR picks between its two schedules: TENSOR_RING
 * (persistent CTAs, shared-memory ring of staged input blocks; needs forward-moving bands
 * of at most 224 features) and TENSOR_TILE (one CTA per tile x slice; any NUM_REGIONS). All are bit-exact;
 * forcing a kernel on an ineligible model makes cloud_eval fail with IDASH_B200_ERR_INVALID. last_kernel()
 * reports what the last launch used (IMAD, TENSOR_TILE or TENSOR_RING). */
enum { IDASH_B200_KERNEL_AUTO = 0, IDASH_B200_KERNEL_IMAD = 1, IDASH_B200_KERNEL_TENSOR = 2,
       IDASH_B200_KERNEL_TENSOR_TILE = 3, IDASH_B200_KERNEL_TENSOR_RING = 4 };
int idash_b200_set_kernel(idash_b200_ctx *ctx, int which);
int idash_b200_last_kernel(const idash_b200_ctx *ctx);
/* pinned host memory for the *_host entry points */
int idash_b200_host_alloc(void **ptr, size_t bytes);
int idash_b200_host_free(void *ptr);
/* page-lock / unlock memory the caller allocated itself (e.g. the slab a ciphertext file was read into) */
int idash_b200_host_register(void *ptr, size_t bytes);
int idash_b200_host_unregister(void *ptr);

/* ---- model: replaces the per-call deep copy of Model (eval/idash.cpp:772) by a one-time compile of
 *      the coefficient maps into a device-resident block-banded layout ---------------------------- */
int idash_b200_model_upload(idash_b200_ctx *ctx, const idash_b200_model_desc *desc, idash_b200_model **model);
/* Uploads a layout that was compiled (idash_b200_layout_compile_ex) or loaded from the model cache (idash_b200_layout_load)
 * beforehand; the model takes ownership of `layout` (also on failure). struct idash_b200_layout is declared in idash_b200_layout.h. */
struct idash_b200_layout;
int idash_b200_model_upload_layout(idash_b200_ctx *ctx, struct idash_b200_layout *layout, idash_b200_model **model);
/* The same model on another device (ctx of that device): the compiled layout is shared, only the device arrays are uploaded again.
 * What a process that shards one evaluation over several GPUs does once per GPU. */
int idash_b200_model_clone(idash_b200_ctx *ctx, const idash_b200_model *model, idash_b200_model **clone);
int idash_b200_model_free(idash_b200_model *model);
int idash_b200_model_get_info(const idash_b200_model *model, idash_b200_model_info *info);

/* ---- cloud_compute_score (eval/idash.cpp:763-848) --------------------------------------------
 * out->count must equal the model's n_rows. Row r of the model (order of desc->out_bidx) is written
 * to output slot slot_of_row[r], or to slot r if slot_of_row is NULL. For RECORDS output the record
 * header {out_bidx, 84} and the variance are written too, so with the reference's record order in
 * slot_of_row the buffer is the image of encrypted_prediction.bin (after its 8-byte count). */
int idash_b200_cloud_eval_host(idash_b200_ctx *ctx, const idash_b200_model *model, const idash_b200_cts *in,
                               const idash_b200_cts *out, const uint32_t *slot_of_row);
int idash_b200_cloud_eval_device(idash_b200_ctx *ctx, const idash_b200_model *model, const idash_b200_cts *in,
                                 const idash_b200_cts *out, const uint32_t *slot_of_row, void *cuda_stream);
/* One GPU's share of an evaluation sharded by contiguous target ranges (the loop being cut is eval/idash.cpp:779-790):
 * model rows [row_begin, row_end), row_begin a multiple of 64 (IDASH_B200_TILE_ROWS) and row_end a multiple of 64 or the model's row
 * count. `out` describes the WHOLE output array (out->count = model rows, caller rows in output-bigIndex order); only the rows of
 * the range are written, only the input ciphertexts their windows touch are copied to the device (PACKED inputs in identity
 * order), and the device staging buffers are sized for the range. One host thread per GPU calls this on its own ctx / model copy. */
int idash_b200_cloud_eval_host_rows(idash_b200_ctx *ctx, const idash_b200_model *model, const idash_b200_cts *in,
                                    const idash_b200_cts *out, uint64_t row_begin, uint64_t row_end);
/* Input ciphertext indices [*ct_begin, *ct_end) that the windows of model rows [row_begin, row_end) touch (same row-range rules as
 * cloud_eval_host_rows): the slab a GPU that owns that target range needs ("derive the slab from the model itself"). */
int idash_b200_model_input_range(const idash_b200_model *model, uint64_t row_begin, uint64_t row_end, uint32_t *ct_begin, uint32_t *ct_end);
/* The same model on n_batches input / output sets (in[b] -> out[b], b < n_batches) in one call: what a GPU that owns a target
 * range does for several sample batches (BASELINE configs[4]). When the ring kernel takes the model, all sets are PACKED
 * with inputs in identity order (index == NULL) and have the same sizes, they are evaluated by ONE launch (n_batches <= 8);
 * otherwise one after another on the same stream. Results are identical to n_batches calls of cloud_eval_device. */
int idash_b200_cloud_eval_device_batched(idash_b200_ctx *ctx, const idash_b200_model *model, uint32_t n_batches,
                                         const idash_b200_cts *in, const idash_b200_cts *out, void *cuda_stream);

/* n models of the same shape (same output rows, NUM_REGIONS, REGION_SIZE; NUM_SAMPLES and coefficients may differ), each on its own
 * input / output set: in[b] -> out[b] under model[b]. BASELINE configs[3] -- the population-stratified model sets, which the reference
 * evaluates as separate `cloud` runs -- in ONE launch of the persistent kernel (n <= 8, PACKED identity inputs of equal size);
 * otherwise one launch per pair on the same stream. Results are identical to n calls of cloud_eval_device. */
int idash_b200_cloud_eval_device_multi_model(idash_b200_ctx *ctx, uint32_t n, const idash_b200_model *const *model,
                                             const idash_b200_cts *in, const idash_b200_cts *out, void *cuda_stream);

/* One evaluation sharded over n_gpus GPUs of this process by contiguous target ranges, with the data resident on the FIRST GPU
 * (scatter of the input slabs and gather of the outputs over NVLink; the loop being cut is eval/idash.cpp:779-790).
 * ctx[g] / model[g]: one context per GPU and the model uploaded to it (model[0] and its idash_b200_model_clone()s). `in` (PACKED, identity
 * order) and `out` (PACKED or RECORDS, rows in output-bigIndex order) are DEVICE memory of ctx[0]'s GPU. GPU g > 0 receives the slab its
 * range reads with a peer copy and its kernels store their rows directly into `out` through the peer mapping (fused gather). Synchronous:
 * returns when every GPU has finished (IDASH_B200_ERR_MISSING_INPUT as for cloud_eval_host). The caller's earlier work on `in` must be complete. */
int idash_b200_cloud_eval_multi_device(uint32_t n_gpus, idash_b200_ctx *const *ctx, const idash_b200_model *const *model,
                                       const idash_b200_cts *in, const idash_b200_cts *out);

/* After a *_device call and a stream synchronise: IDASH_B200_ERR_MISSING_INPUT if a kernel met a model
 * entry whose ciphertext was not supplied, else IDASH_B200_OK. Clears the flag. */
int idash_b200_check_device_status(idash_b200_ctx *ctx);

/* ---- decrypt_predictions (eval/idash.cpp:681-761): phase = b - key*a mod (X^1024+1, 2^32), exact;
 *      scores[i][j] = (float)((double)(int32)phase[i][j] / 2^32) for j < num_samples ---------------
 * key: 1024 int32 in {0,1} (TLweKey::key[0].coefs, tfhe_io.cpp:409-416). scores: [count][num_samples]
 * floats. phase: optional [count][1024] words (NULL to skip). */
int idash_b200_decrypt_host(idash_b200_ctx *ctx, const int32_t *key, uint32_t num_samples, const idash_b200_cts *in,
                            float *scores, uint32_t *phase);
int idash_b200_decrypt_device(idash_b200_ctx *ctx, const int32_t *key_host, uint32_t num_samples,
                              const idash_b200_cts *in, float *scores, uint32_t *phase, void *cuda_stream);
/* Which kernel idash_b200_decrypt_* launches. TENSOR: the negacyclic product as a tcgen05 int8 GEMM of the key's
 * 1024 x 1024 Toeplitz matrix against the byte planes of a (32 ciphertexts per group), one CTA per SM. TENSOR_PAIR
 * (= AUTO): the same GEMM on CTA pairs (tcgen05.mma.cta_group::2, M = 256; each CTA holds half of a group's planes, two
 * groups resident). IADD = the CUDA-core kernel (one CTA per ciphertext, ~512 integer adds per phase coefficient).
 * All three are exact and bit-identical. */
enum { IDASH_B200_DECRYPT_AUTO = 0, IDASH_B200_DECRYPT_IADD = 1, IDASH_B200_DECRYPT_TENSOR = 2, IDASH_B200_DECRYPT_TENSOR_PAIR = 3 };
int idash_b200_set_decrypt_kernel(idash_b200_ctx *ctx, int which);
int idash_b200_last_decrypt_kernel(const idash_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* IDASH_B200_H */
