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

typedef struct idash_b200_layout idash_b200_layout;

int idash_b200_layout_compile(const idash_b200_model_desc *desc, idash_b200_layout **layout);
int idash_b200_layout_free(idash_b200_layout *layout);
int idash_b200_layout_get_info(const idash_b200_layout *layout, idash_b200_model_info *info);
const idash_b200_group *idash_b200_layout_groups(const idash_b200_layout *layout, uint64_t *n_groups);
const idash_b200_entry *idash_b200_layout_entries(const idash_b200_layout *layout, uint64_t *n_entries);
/* variance CSR: var_ptr[n_rows+1] (uint64), var_ct[nv] (uint32), var_w[nv] (double) */
const uint64_t *idash_b200_layout_var_ptr(const idash_b200_layout *layout);
const uint32_t *idash_b200_layout_var_ct(const idash_b200_layout *layout, uint64_t *n);
const double *idash_b200_layout_var_w(const idash_b200_layout *layout);
/* output bigIndex per caller row (copy of desc->out_bidx) */
const uint32_t *idash_b200_layout_out_bidx(const idash_b200_layout *layout);

#ifdef __cplusplus
}
#endif
#endif
