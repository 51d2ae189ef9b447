// host_selftest -- exercises the file-format half of the host layer without a GPU (driven by tests/test_host.py):
//   host_selftest <dir with params.bin keys.bin encrypted_data.bin encrypted_prediction.bin model/> <out dir>
// writes into <out dir>: params.txt, model.txt, key.txt, encrypted_data.rt.bin, encrypted_prediction.rt.bin, result.csv,
// result_bypos.csv (scores are a deterministic function of position / variant / sample, not a decryption).
#include "idash_host.h"
#include "parse_vw.h"

#include <algorithm>
#include <charconv>
#include <cstring>
#include <cstdio>
#include <fstream>
#include <map>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: host_selftest <in dir> <out dir>\n"); return 2; }
    const std::string in = argv[1], out = argv[2];
    IdashParams params;
    read_params(params, in + "/" PARAMS_FILE);
    {
        std::ofstream f(out + "/params.txt");
        f << params.NUM_SAMPLES << " " << params.NUM_INPUT_POSITIONS << " " << params.NUM_OUTPUT_POSITIONS << " " << params.NUM_INPUT_FEATURES
          << " " << params.NUM_OUTPUT_FEATURES << " " << params.NUM_REGIONS << " " << params.REGION_SIZE << "\n";
        for (const auto &pn : params.out_position_names)
            f << pn.first << " " << pn.second << " " << params.outBigIdx(pn.first, 0) << " " << params.outBigIdx(pn.first, 1) << " "
              << params.outBigIdx(pn.first, 2) << "\n";
    }
    Model model;
    read_model(model, params, in + "/model");
    {
        std::ofstream f(out + "/model.txt");
        std::map<FeatBigIndex, std::map<FeatBigIndex, int32_t>> sorted;
        for (const auto &r : model.model) sorted[r.first].insert(r.second.begin(), r.second.end());
        for (const auto &r : sorted)
            for (const auto &c : r.second) f << r.first << " " << c.first << " " << c.second << "\n";
    }
    IdashKey key;
    read_key(key, in + "/" KEYS_FILE);
    {
        std::ofstream f(out + "/key.txt");
        f << key.idashParams->NUM_SAMPLES << " " << key.tlweKey->params->N << " " << key.tlweKey->params->k << "\n";
        for (int i = 0; i < key.tlweKey->params->N; ++i) f << key.tlweKey->key[0].coefs[i] << "\n";
    }
    EncryptedData enc;
    read_encrypted_data(enc, params, in + "/" ENCRYPTED_DATA_FILE);
    write_encrypted_data(enc, params, out + "/encrypted_data.rt.bin");
    EncryptedPredictions pred;
    read_encrypted_predictions(pred, params, in + "/" ENCRYPTED_PREDICTION_FILE);
    // touch one variance through the TLweSample view: the writer must pick it up
    if (!pred.score.empty()) pred.score.begin()->second->current_variance += 0.0;
    write_encrypted_predictions(pred, params, out + "/encrypted_prediction.rt.bin");

    DecryptedPredictions dec;
    for (const auto &it : params.out_features_index)
        for (int snp = 0; snp < 3; ++snp) {
            auto &v = dec.score[it.first][snp];
            v.resize(params.NUM_SAMPLES);
            for (uint32_t s = 0; s < params.NUM_SAMPLES; ++s)
                v[s] = (float) ((double) (int32_t) ((uint32_t) it.first * 2654435761u + (uint32_t) snp * 40503u + s * 2246822519u) / 4294967296.0);
        }
    write_decrypted_predictions(dec, params, out + "/" RESULT_FILE, true);
    write_decrypted_predictions(dec, params, out + "/" RESULT_BYPOS_FILE, false);
    // the csv writer formats with std::to_chars(general, 6); the reference streams `ostream << float` = printf("%g"):
    // cross-check the two on two million floats (scores in [-0.5, 0.5), random bit patterns, edge values)
    {
        uint64_t x = 88172645463325252ull;
        char a[64], b[64];
        for (int i = 0; i < 2000000; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            float f;
            if (i % 3 == 0) f = (float) ((double) (int32_t) (uint32_t) x / 4294967296.0);
            else if (i % 3 == 1) { uint32_t u = (uint32_t) (x >> 16); memcpy(&f, &u, 4); if (f != f || f - f != 0.0f) continue; }
            else f = (float) ((double) (int32_t) (uint32_t) (x >> 40) / 4294967296.0);
            if (i < 8) { const float edge[8] = {0.0f, -0.0f, 1.0f, -0.5f, 1e-5f, 9.99999e-5f, 123456.0f, 1234567.0f}; f = edge[i]; }
            const int n = snprintf(a, sizeof(a), "%g", (double) f);
            const auto r = std::to_chars(b, b + sizeof(b), (double) f, std::chars_format::general, 6);
            if ((size_t) n != (size_t) (r.ptr - b) || memcmp(a, b, (size_t) n) != 0) {
                *r.ptr = 0;
                fprintf(stderr, "float formatting differs: printf '%s' to_chars '%s'\n", a, b);
                return 1;
            }
        }
    }
    printf("host_selftest ok: %zu model rows, %zu input cts, %zu prediction cts\n", model.model.size(), enc.enc_data.size(), pred.score.size());
    return 0;
}
