// The reference-side binding of INTEGRATION.md section B as a real translation unit: new bodies for cloud_compute_score
// (eval/idash.cpp:763-848) and decrypt_predictions (eval/idash.cpp:681-761) written against the REFERENCE's own headers (eval/idash.h,
// <tfhe.h>) that call the C ABI of include/idash_b200.h. A maintainer drops this file into eval/, removes the two bodies from
// eval/idash.cpp and links libidash_b200. tests/test_host.py::test_reference_glue_compiles_against_reference_headers type-checks it
// against /root/reference where that tree exists, so the snippet in INTEGRATION.md (generated from this file) cannot rot.
// eval/idash_b200_glue.cpp  -- add to eval/CMakeLists.txt:  target_link_libraries(idash idash_b200)
#include <cstring>
#include <unordered_map>
#include <vector>

#include "idash.h"
#include "idash_b200.h"

static idash_b200_ctx *ctx() {
    static idash_b200_ctx *c = nullptr;
    if (!c && idash_b200_init(&c, /*device*/ 0) != IDASH_B200_OK) DIE_DRAMATICALLY(idash_b200_last_error());
    return c;
}

void cloud_compute_score(EncryptedPredictions &enc_preds, const EncryptedData &enc_data, const Model &model,
                         const IdashParams &params) {
    REQUIRE_DRAMATICALLY(params.k == 1, "blah");
    // 1. Model -> CSR in the map's iteration order (the order eval/idash.cpp:772-777 walks it)
    std::vector<uint32_t> out_bidx, col; std::vector<uint64_t> row_ptr{0}; std::vector<int32_t> coef;
    for (const auto &row : model.model) {
        out_bidx.push_back(row.first);
        for (const auto &c : row.second) {
            if (c.first != params.constant_bigIndex()) (void) enc_data.getTLWE(c.first, params);   // "shit happens before"
            col.push_back(c.first); coef.push_back(c.second);
        }
        row_ptr.push_back(col.size());
    }
    idash_b200_model_desc d{params.NUM_SAMPLES, params.NUM_REGIONS, params.REGION_SIZE, out_bidx.size(),
                            out_bidx.data(), row_ptr.data(), col.data(), coef.data()};
    idash_b200_model *m = nullptr;
    if (idash_b200_model_upload(ctx(), &d, &m)) DIE_DRAMATICALLY(idash_b200_last_error());
    // 2. inputs: gather the TLweSample polynomials into one pinned packed array (index + variance alongside)
    const size_t n_in = enc_data.enc_data.size();
    uint32_t *in_words, *in_idx; double *in_var;
    idash_b200_host_alloc((void **) &in_words, n_in * IDASH_B200_CT_BYTES);
    idash_b200_host_alloc((void **) &in_idx, n_in * 4); idash_b200_host_alloc((void **) &in_var, n_in * 8);
    size_t i = 0;
    for (const auto &it : enc_data.enc_data) {
        memcpy(in_words + i * 2048, it.second->a[0].coefsT, 4096);
        memcpy(in_words + i * 2048 + 1024, it.second->a[1].coefsT, 4096);
        in_idx[i] = it.first; in_var[i] = it.second->current_variance; ++i;
    }
    // 3. outputs: created in the reference's order (createAndGet), filled from one packed result array
    uint32_t *out_words; double *out_var;
    idash_b200_host_alloc((void **) &out_words, out_bidx.size() * IDASH_B200_CT_BYTES);
    idash_b200_host_alloc((void **) &out_var, out_bidx.size() * 8);
    idash_b200_cts in{IDASH_B200_LAYOUT_PACKED, in_words, n_in, in_idx, in_var};
    idash_b200_cts out{IDASH_B200_LAYOUT_PACKED, out_words, out_bidx.size(), nullptr, out_var};
    if (idash_b200_cloud_eval_host(ctx(), m, &in, &out, nullptr)) DIE_DRAMATICALLY(idash_b200_last_error());
    for (size_t r = 0; r < out_bidx.size(); ++r) {
        TLweSample *s = enc_preds.createAndGet(out_bidx[r], params.tlweParams);
        memcpy(s->a[0].coefsT, out_words + r * 2048, 4096);
        memcpy(s->a[1].coefsT, out_words + r * 2048 + 1024, 4096);
        s->current_variance = out_var[r];
    }
    idash_b200_model_free(m);
    idash_b200_host_free(in_words); idash_b200_host_free(in_idx); idash_b200_host_free(in_var);
    idash_b200_host_free(out_words); idash_b200_host_free(out_var);
}

void decrypt_predictions(DecryptedPredictions &predictions, const EncryptedPredictions &enc_preds, const IdashKey &key) {
    const IdashParams &params = *key.idashParams;
    const size_t n = enc_preds.score.size(); const uint32_t S = params.NUM_SAMPLES;
    uint32_t *words; float *scores; std::vector<uint32_t> idx; idx.reserve(n);
    idash_b200_host_alloc((void **) &words, n * IDASH_B200_CT_BYTES);
    idash_b200_host_alloc((void **) &scores, n * S * sizeof(float));
    size_t i = 0;
    for (const auto &it : enc_preds.score) {
        memcpy(words + i * 2048, it.second->a[0].coefsT, 4096);
        memcpy(words + i * 2048 + 1024, it.second->a[1].coefsT, 4096);
        idx.push_back(it.first); ++i;
    }
    idash_b200_cts in{IDASH_B200_LAYOUT_PACKED, words, n, nullptr, nullptr};
    if (idash_b200_decrypt_host(ctx(), key.tlweKey->key[0].coefs, S, &in, scores, nullptr)) DIE_DRAMATICALLY(idash_b200_last_error());
    std::unordered_map<FeatBigIndex, size_t> slot;
    for (size_t k = 0; k < n; ++k) slot[idx[k]] = k;
    for (const auto &it : params.out_features_index)
        for (int snp = 0; snp < 3; ++snp) {
            const float *src = scores + slot.at(it.second[snp]) * S;
            predictions.score[it.first][snp].assign(src, src + S);
        }
    idash_b200_host_free(words); idash_b200_host_free(scores);
}
