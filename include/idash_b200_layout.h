/* idash_b200_layout.h -- the device-resident "block-banded" model layout that
 * idash_b200_model_upload() builds from the reference's Model (eval/idash.h:129-134), exposed so
 * that host code (parse_vw-side loaders, tests) can compile, inspect and cache it without a GPU.
 *
 * Rows are sorted by output bigIndex and cut into GROUPS of two consecutive target SNPs (A and B) x
 * their (up to) three one-hot variants = 6 rows. The three rows of a target share one window of
 * neighbouring tag-SNP ciphertexts and consecutive targets share most of it, so a group lists each
 * distinct (input ciphertext, rotation) pair ONCE, in three runs:
 *     [ nA entries used only by target A | nAB entries used by both | nB entries only by B ]
 * and every entry carries the 6 integer coefficients (0 where a row does not use it) -- a dense
 * 6 x (nA+nAB+nB) coefficient block per group. The kernel loads each input word once per group and
 * issues 3 (A-only / B-only) or 6 (shared) multiply-adds with it.
 *
 * An input feature f (bigIndex) lives in ciphertext f / NR, region f % NR (eval/idash.h:87-89); the
 * per-region temporaries + torusPolynomialMulByXai(2N - r*RS) of eval/idash.cpp:795-836 are folded
 * into `shift = r*RS`: the kernel reads word (i + shift) mod 1024 of the input polynomial, negated
 * when i + shift >= 1024.
 */
#ifndef IDASH_B200_LAYOUT_H
#define IDASH_B200_LAYOUT_H

#include "idash_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IDASH_B200_GROUP_ROWS 6u
#define IDASH_B200_NO_ROW 0xFFFFFFFFu

/* 32 bytes, 16-byte aligned: read by the kernel as two 128-bit uniform loads */
typedef struct idash_b200_entry {
    uint32_t ct;      /* input ciphertext index (bigIndex / NR) */
    uint32_t shift;   /* (bigIndex % NR) * RS, < 1024 */
    int32_t coef[6];  /* rows A0 A1 A2 B0 B1 B2 */
} idash_b200_entry;

/* 64 bytes */
typedef struct idash_b200_group {
    uint32_t entry_begin;  /* first entry of the group in the entry array */
    uint32_t n_a, n_ab, n_b;
    uint32_t row[6];       /* caller row number (position in desc->out_bidx) or IDASH_B200_NO_ROW */
    int32_t bias[6];       /* "Constant" coefficient; the kernel adds bias * 2^18 to b[0..S) */
} idash_b200_group;

/* per-row variance bookkeeping (tlwe-functions.cpp:175; only region 0 counts):
 * var_out[row] = sum over e in [var_ptr[row], var_ptr[row+1]) of var_w[e] * var_in[var_ct[e]],
 * var_w = (double)(int32)(coef*coef). Stored as CSR over caller rows. */

/* ---- band tiles: the operand layout of the tcgen05 (int8 tensor-core) kernel ----------------------
 * Rows sorted by output bigIndex are cut into TILES of IDASH_B200_TILE_ROWS consecutive rows. A tile's
 * rows only use input features (bigIndex) inside one contiguous band [f_base, f_base + K), K a multiple
 * of 32, so the tile is a dense (rows x K) coefficient block times the (K x 2048 words) block of
 * (rotated) input ciphertexts:  out[row][word] = sum_k coef[row][k] * X[f_base + k][word]  mod 2^32.
 * The kernel evaluates it with 8-bit limbs on the tensor cores: X = sum_j 2^(8j) X_j (4 unsigned limbs, split
 * on the fly), coef = c_lo + 256 * c_hi with BALANCED signed limbs c_lo, c_hi in [-128, 127] (so
 * IDASH_B200_TILE_COEF_MIN <= coef <= IDASH_B200_TILE_COEF_MAX),
 *     out = sum_{w=0..3} 2^(8w) * P_w,    P_w = X_w * c_lo + X_(w-1) * c_hi     (int32 accumulators).
 * The coefficient image of a tile is stored ready to be copied into shared memory as the K-major,
 * no-swizzle tcgen05 operand, one 4096-byte CHUNK per 32 features (= one MMA K step) at
 * b_off + 4096 * (k / 32). A chunk is two 16-feature halves of 2048 bytes; a half is the 128-row operand
 * [c_lo of rows 0..63 | c_hi of rows 0..63], 16 bytes per row: byte
 *     ((k % 32) / 16) * 2048 + limb * 1024 + n * 16 + (k % 16)
 * is limb (0 = c_lo, 1 = c_hi) of coef[row n][feature f_base + k]. One N = 128 MMA of limb plane X_j against a
 * chunk therefore adds X_j c_lo to P_j and X_j c_hi to P_(j+1), which sit in adjacent TMEM columns.
 * The "Constant" (bias) is NOT part of the band: the kernel adds bias * 2^18 to b[0..S) in its epilogue. */
#define IDASH_B200_TILE_ROWS 64u
#define IDASH_B200_TILE_COEF_MIN (-32896)
#define IDASH_B200_TILE_COEF_MAX 32639
#define IDASH_B200_TILE_KMAX 256u   /* widest band (features) a tile may have; wider models use the IMAD kernel */
/* Every band starts on a multiple of 32 features (a "block"). When, in addition, bands
 * only move forward from tile to tile and are at most RING_KMAX wide, the persistent kernel keeps the staged
 * blocks of consecutive tiles in a shared-memory ring and stages every input block once per CTA. */
#define IDASH_B200_RING_KMAX 224u

typedef struct idash_b200_tile {
    uint32_t f_base;    /* first input bigIndex of the band */
    uint32_t K;         /* band width in features, multiple of 32, <= IDASH_B200_TILE_KMAX */
    uint64_t b_off;     /* byte offset of the tile's coefficient image (16-byte aligned) */
    uint32_t used_off;  /* offset (uint32 words) of the K/32-word mask of features with a non-zero coefficient */
    uint32_t n_valid;   /* row slots of the tile (the last tile may be partial); a slot whose row was evicted holds NO_ROW */
    uint32_t flags;     /* bit 0: full tile whose caller rows are consecutive integers starting at tile_rows[64 t] */
    uint32_t pad;
} idash_b200_tile;      /* 32 bytes */

typedef struct idash_b200_layout idash_b200_layout;

/* Compiles the model. Every pass is a parallel loop over rows or tiles (IDASH_B200_THREADS host threads, default: all).
 * A row that no tile can hold -- a coefficient outside [TILE_COEF_MIN, TILE_COEF_MAX], or a window that does not fit into its tile's
 * band of at most RING_KMAX features -- is EVICTED from its tile (tile_rows = NO_ROW there) and listed in the OVERFLOW groups, which
 * the IMAD kernel evaluates beside the tensor-core kernel: one outlier window never changes the kernel for the other rows.
 * flags: IDASH_B200_COMPILE_GROUPS_ALL also builds the IMAD groups of EVERY row (what idash_b200_layout_groups returns and what
 * forcing IDASH_B200_KERNEL_IMAD needs; built on demand otherwise). idash_b200_layout_compile = compile_ex with that flag. */
#define IDASH_B200_COMPILE_DEFAULT 0u
#define IDASH_B200_COMPILE_GROUPS_ALL 1u
int idash_b200_layout_compile_ex(const idash_b200_model_desc *desc, uint32_t flags, idash_b200_layout **layout);
int idash_b200_layout_compile(const idash_b200_model_desc *desc, idash_b200_layout **layout);
int idash_b200_layout_ensure_groups_all(idash_b200_layout *layout);
int idash_b200_layout_free(idash_b200_layout *layout);
/* The cached packed model ("models.bin"): the compiled layout as one file, so that a later run skips the 3 G
 * .hr text files (eval/parse_vw.cpp:8-30, eval/idash.cpp:66-90) and the compile. `key` is the caller's fingerprint of what the model
 * was compiled from (the host layer hashes params.bin and the model directory listing); load fails with IDASH_B200_ERR_INVALID when
 * the file is missing, truncated, from another library version or saved under another key. */
int idash_b200_layout_save(const idash_b200_layout *layout, const char *path, uint64_t key);
int idash_b200_layout_load(const char *path, uint64_t key, idash_b200_layout **layout);
/* the model itself as the layout keeps it (per caller row: Constant + entries sorted by input bigIndex) */
const uint64_t *idash_b200_layout_feat_ptr(const idash_b200_layout *layout);
const uint32_t *idash_b200_layout_feat_bidx(const idash_b200_layout *layout, uint64_t *n);
const int32_t *idash_b200_layout_feat_coef(const idash_b200_layout *layout);
const int32_t *idash_b200_layout_bias(const idash_b200_layout *layout);
/* IMAD groups of the overflow rows only */
const idash_b200_group *idash_b200_layout_overflow_groups(const idash_b200_layout *layout, uint64_t *n_groups);
const idash_b200_entry *idash_b200_layout_overflow_entries(const idash_b200_layout *layout, uint64_t *n_entries);
int idash_b200_layout_get_info(const idash_b200_layout *layout, idash_b200_model_info *info);
/* IMAD groups of every row (*n = 0 unless compiled with COMPILE_GROUPS_ALL / after layout_ensure_groups_all) */
const idash_b200_group *idash_b200_layout_groups(const idash_b200_layout *layout, uint64_t *n_groups);
const idash_b200_entry *idash_b200_layout_entries(const idash_b200_layout *layout, uint64_t *n_entries);
/* variance CSR: var_ptr[n_rows+1] (uint64), var_ct[nv] (uint32), var_w[nv] (double) */
const uint64_t *idash_b200_layout_var_ptr(const idash_b200_layout *layout);
const uint32_t *idash_b200_layout_var_ct(const idash_b200_layout *layout, uint64_t *n);
const double *idash_b200_layout_var_w(const idash_b200_layout *layout);
/* output bigIndex per caller row (copy of desc->out_bidx) */
const uint32_t *idash_b200_layout_out_bidx(const idash_b200_layout *layout);
/* band tiles. *n_tiles = 0 when no row is eligible for the tensor-core kernels. tile_rows / tile_bias: [n_tiles * TILE_ROWS]
 * caller row (IDASH_B200_NO_ROW = padding or an evicted row) and Constant of every tile row. */
const idash_b200_tile *idash_b200_layout_tiles(const idash_b200_layout *layout, uint64_t *n_tiles);
const uint32_t *idash_b200_layout_tile_rows(const idash_b200_layout *layout);
const int32_t *idash_b200_layout_tile_bias(const idash_b200_layout *layout);
const uint8_t *idash_b200_layout_tile_coef(const idash_b200_layout *layout, uint64_t *n_bytes);
const uint32_t *idash_b200_layout_tile_used(const idash_b200_layout *layout, uint64_t *n_words);
/* ring variant: bit f of the mask = some row has a non-zero coefficient on input feature f (*n_words = 0: not eligible) */
const uint32_t *idash_b200_layout_feat_used(const idash_b200_layout *layout, uint64_t *n_words);

#ifdef __cplusplus
}
#endif
#endif
