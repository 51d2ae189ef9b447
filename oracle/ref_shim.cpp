// TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from the product path).
//
// A C-ABI shim that lets tests/ and bench.py's reference arm call the UNMODIFIED reference
// functions (compiled from /root/reference by oracle/build_ref.sh) on flat arrays:
//   cloud_compute_score   eval/idash.cpp:763-848   (declared eval/idash.h:251-252)
//   decrypt_predictions   eval/idash.cpp:681-761   (declared eval/idash.h:254)
//   read_model            eval/idash.cpp:66-90     (declared eval/idash.h:229)
//   read_params           eval/idash.cpp:264-272
//   torusPolynomialSubMulRKaratsuba  tfhe/src/libtfhe/multiplication.cpp:162-177 (exact product)
// The shim only marshals data into / out of the reference's own containers (eval/idash.h:45-219);
// it contains no arithmetic of its own.
#include "idash.h"

#include <algorithm>
#include <cstring>
#include <vector>
#include <sys/time.h>

static double wall() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec;
}

static void fill_params(IdashParams &p, uint32_t S, uint32_t NR, uint32_t RS) {
    p.NUM_SAMPLES = S;
    p.NUM_REGIONS = NR;
    p.REGION_SIZE = RS;
    p.NUM_INPUT_POSITIONS = p.NUM_OUTPUT_POSITIONS = 0;
    p.NUM_INPUT_FEATURES = p.NUM_OUTPUT_FEATURES = 0;
}

extern "C" {

// Runs the reference cloud_compute_score.
//   in_idx[n_in]          ciphertext index (FeatIndex = bigIndex / NR) of every input ciphertext
//   in_ct[n_in][2048]     words a[0..1024) then b[0..1024)
//   in_var[n_in]          current_variance
//   model in CSR form: out_bidx[n_out], row_ptr[n_out+1], col[nnz] (input bigIndex, 0xFFFFFFFF =
//   Constant), coef[nnz]
//   out_ct[n_out][2048], out_var[n_out]  in the order of out_bidx[]
// Returns the wall time of the cloud_compute_score call alone (what cloud.cpp:16-18 prints as
// "fhe wall time").
double ref_cloud_compute_score(uint32_t S, uint32_t NR, uint32_t RS,
                               uint64_t n_in, const uint32_t *in_idx, const uint32_t *in_ct,
                               const double *in_var,
                               uint64_t n_out, const uint32_t *out_bidx, const uint64_t *row_ptr,
                               const uint32_t *col, const int32_t *coef,
                               uint32_t *out_ct, double *out_var) {
    IdashParams params;
    fill_params(params, S, NR, RS);
    const TLweParams *tp = params.tlweParams;
    const uint32_t N = params.N;

    EncryptedData enc;
    TLweSample *pool = new_TLweSample_array((int32_t) n_in, tp);
    for (uint64_t i = 0; i < n_in; ++i) {
        memcpy(pool[i].a[0].coefsT, in_ct + i * 2 * N, sizeof(uint32_t) * N);
        memcpy(pool[i].a[1].coefsT, in_ct + i * 2 * N + N, sizeof(uint32_t) * N);
        pool[i].current_variance = in_var ? in_var[i] : 0.;
        enc.enc_data.emplace(in_idx[i], &pool[i]);
    }
    Model model;
    for (uint64_t o = 0; o < n_out; ++o) {
        auto &row = model.model[out_bidx[o]];
        for (uint64_t e = row_ptr[o]; e < row_ptr[o + 1]; ++e) row[col[e]] = coef[e];
    }
    EncryptedPredictions preds;
    const double t0 = wall();
    cloud_compute_score(preds, enc, model, params);
    const double t1 = wall();
    for (uint64_t o = 0; o < n_out; ++o) {
        TLweSample *s = preds.score.at(out_bidx[o]);
        if (out_ct) {
            memcpy(out_ct + o * 2 * N, s->a[0].coefsT, sizeof(uint32_t) * N);
            memcpy(out_ct + o * 2 * N + N, s->a[1].coefsT, sizeof(uint32_t) * N);
        }
        if (out_var) out_var[o] = s->current_variance;
        delete_TLweSample(s);
    }
    delete_TLweSample_array((int32_t) n_in, pool);
    return t1 - t0;
}

// Runs the reference decrypt_predictions (double-precision FFT phase, eval/idash.cpp:681-761).
// n_ct must be a multiple of 3 (ciphertext 3p+v = position p, variant v).
//   key[1024] in {0,1}; ct[n_ct][2048]; scores[n_ct][S] floats.
// Returns the wall time of the decrypt_predictions call alone (decrypt.cpp:21-23).
double ref_decrypt_predictions(uint32_t S, const int32_t *key, uint64_t n_ct, const uint32_t *ct,
                               float *scores) {
    IdashParams *params = new IdashParams();
    fill_params(*params, S, 1024 / S, 1024 / (1024 / S));
    const TLweParams *tp = params->tlweParams;
    const uint32_t N = params->N;
    TLweKey *tk = new_TLweKey(tp);
    memcpy(tk->key[0].coefs, key, sizeof(int32_t) * N);
    IdashKey ikey(params, tk);
    const uint64_t n_pos = n_ct / 3;
    for (uint64_t p = 0; p < n_pos; ++p)
        for (int v = 0; v < 3; ++v) params->registerOutBigIdx(p, v, (FeatBigIndex) (3 * p + v));
    EncryptedPredictions preds;
    TLweSample *pool = new_TLweSample_array((int32_t) (3 * n_pos), tp);
    for (uint64_t i = 0; i < 3 * n_pos; ++i) {
        memcpy(pool[i].a[0].coefsT, ct + i * 2 * N, sizeof(uint32_t) * N);
        memcpy(pool[i].a[1].coefsT, ct + i * 2 * N + N, sizeof(uint32_t) * N);
        preds.score.emplace((FeatBigIndex) i, &pool[i]);
    }
    DecryptedPredictions dec;
    const double t0 = wall();
    decrypt_predictions(dec, preds, ikey);
    const double t1 = wall();
    for (uint64_t p = 0; p < n_pos; ++p)
        for (int v = 0; v < 3; ++v)
            memcpy(scores + (3 * p + v) * S, dec.score.at(p)[v].data(), sizeof(float) * S);
    delete_TLweSample_array((int32_t) (3 * n_pos), pool);
    delete_TLweKey(tk);
    delete params;
    return t1 - t0;
}

// phase[n_ct][1024] = b - key*a with the reference's own FFT tLwePhase (use_fft != 0,
// tlwe-functions.cpp:64-71) or with TFHE's exact Karatsuba product (multiplication.cpp:162-177).
void ref_tlwe_phase(const int32_t *key, uint64_t n_ct, const uint32_t *ct, int use_fft,
                    uint32_t *phase) {
    const TLweParams *tp = IdashParams::tlweParams;
    const uint32_t N = IdashParams::N;
    TLweKey *tk = new_TLweKey(tp);
    memcpy(tk->key[0].coefs, key, sizeof(int32_t) * N);
    TLweSample *s = new_TLweSample(tp);
    TorusPolynomial *ph = new_TorusPolynomial(N);
    for (uint64_t i = 0; i < n_ct; ++i) {
        memcpy(s->a[0].coefsT, ct + i * 2 * N, sizeof(uint32_t) * N);
        memcpy(s->a[1].coefsT, ct + i * 2 * N + N, sizeof(uint32_t) * N);
        if (use_fft) {
            tLwePhase(ph, s, tk);
        } else {
            torusPolynomialCopy(ph, s->b);
            torusPolynomialSubMulRKaratsuba(ph, &tk->key[0], &s->a[0]);
        }
        memcpy(phase + i * N, ph->coefsT, sizeof(uint32_t) * N);
    }
    delete_TorusPolynomial(ph);
    delete_TLweSample(s);
    delete_TLweKey(tk);
}

// ---- reference model loader (read_params + read_model), exported as sorted CSR -------------
struct RefModel {
    IdashParams params;
    std::vector<uint32_t> out_bidx;
    std::vector<uint64_t> row_ptr;
    std::vector<uint32_t> col;
    std::vector<int32_t> coef;
};

void *ref_model_load(const char *params_file, const char *model_dir) {
    RefModel *m = new RefModel();
    read_params(m->params, params_file);
    Model model;
    read_model(model, m->params, model_dir);
    for (const auto &row : model.model) m->out_bidx.push_back(row.first);
    std::sort(m->out_bidx.begin(), m->out_bidx.end());
    m->row_ptr.push_back(0);
    for (uint32_t o : m->out_bidx) {
        std::vector<std::pair<uint32_t, int32_t>> es(model.model.at(o).begin(), model.model.at(o).end());
        std::sort(es.begin(), es.end());
        for (const auto &e : es) {
            m->col.push_back(e.first);
            m->coef.push_back(e.second);
        }
        m->row_ptr.push_back(m->col.size());
    }
    return m;
}
uint64_t ref_model_rows(void *h) { return ((RefModel *) h)->out_bidx.size(); }
uint64_t ref_model_nnz(void *h) { return ((RefModel *) h)->col.size(); }
void ref_model_geometry(void *h, uint32_t *g7) {
    const IdashParams &p = ((RefModel *) h)->params;
    g7[0] = p.NUM_SAMPLES; g7[1] = p.NUM_INPUT_POSITIONS; g7[2] = p.NUM_OUTPUT_POSITIONS;
    g7[3] = p.NUM_INPUT_FEATURES; g7[4] = p.NUM_OUTPUT_FEATURES; g7[5] = p.NUM_REGIONS;
    g7[6] = p.REGION_SIZE;
}
void ref_model_export(void *h, uint32_t *out_bidx, uint64_t *row_ptr, uint32_t *col, int32_t *coef) {
    RefModel *m = (RefModel *) h;
    memcpy(out_bidx, m->out_bidx.data(), 4 * m->out_bidx.size());
    memcpy(row_ptr, m->row_ptr.data(), 8 * m->row_ptr.size());
    memcpy(col, m->col.data(), 4 * m->col.size());
    memcpy(coef, m->coef.data(), 4 * m->coef.size());
}
void ref_model_free(void *h) { delete (RefModel *) h; }

// constants computed by the reference (eval/idash.cpp:39-45)
void ref_constants(int32_t *c4) {
    c4[0] = IdashParams::ONE_IN_T32; c4[1] = IdashParams::NAN_0_IN_T32;
    c4[2] = IdashParams::NAN_1_IN_T32; c4[3] = IdashParams::NAN_2_IN_T32;
}
int ref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
