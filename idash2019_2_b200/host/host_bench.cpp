// host_bench -- wall time of every file-level phase of the host layer on a pipeline directory, without a GPU:
//   host_bench <dir with params.bin keys.bin encrypted_data.bin encrypted_prediction.bin> <model dir> <out dir> [csv]
// Prints one "phase seconds" line per phase (the phases SURVEY.md 8(f) ranks: model loading, ciphertext file I/O, the
// result CSV writer). The compute stages are not run; scores for the CSV are synthetic.
#include "idash_host.h"

#include <chrono>
#include <cstdio>
#include <cstring>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define PHASE(name, stmt)                                   \
    do {                                                    \
        const double t0_ = now();                           \
        stmt;                                               \
        printf("%-28s %8.3f s\n", name, now() - t0_);       \
        fflush(stdout);                                     \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: host_bench <pipeline dir> <model dir> <out dir> [csv]\n"); return 2; }
    const std::string in = argv[1], model_dir = argv[2], out = argv[3];
    const bool csv = argc >= 5;
    IdashParams params;
    Model model;
    EncryptedData enc;
    EncryptedPredictions pred;
    PHASE("read_params", read_params(params, in + "/" PARAMS_FILE));
    PHASE("read_model", read_model(model, params, model_dir));
    PHASE("read_encrypted_data", read_encrypted_data(enc, params, in + "/" ENCRYPTED_DATA_FILE));
    PHASE("read_encrypted_predictions", read_encrypted_predictions(pred, params, in + "/" ENCRYPTED_PREDICTION_FILE));
    PHASE("write (generic, map order)", write_encrypted_predictions(pred, params, out + "/" ENCRYPTED_PREDICTION_FILE));
    {
        // the cloud stage's case: a slab laid out in the map's iteration order, so that its image is the file
        EncryptedPredictions p2;
        for (const auto &it : pred.score) p2.score.emplace(it.first, nullptr);
        p2.slab = std::make_shared<CtSlab>(pred.score.size());
        uint64_t k = 0;
        for (auto &it : p2.score) {
            it.second = &p2.slab->samples[k];
            memcpy(p2.slab->record(k), pred.slab->record((uint64_t) pred.slab->slot_of(pred.score.at(it.first))), 8208);
            it.second->current_variance = pred.score.at(it.first)->current_variance;
            ++k;
        }
        PHASE("write (slab image)", write_encrypted_predictions(p2, params, out + "/" ENCRYPTED_PREDICTION_FILE));
        EncryptedPredictions p3;
        PHASE("read back", read_encrypted_predictions(p3, params, out + "/" ENCRYPTED_PREDICTION_FILE));
        const bool same = p3.slab->count == p2.slab->count && memcmp(p3.slab->image(), p2.slab->image(), p2.slab->image_bytes()) == 0;
        printf("slab image round trip: %s\n", same ? "identical" : "DIFFERENT");
        if (!same) return 1;
    }
    if (csv) {
        DecryptedPredictions dec;
        PHASE("fill synthetic scores", {
            for (const auto &it : params.out_features_index)
                for (int snp = 0; snp < 3; ++snp) {
                    auto &v = dec.score[it.first][snp];
                    v.resize(params.NUM_SAMPLES);
                    for (uint32_t s = 0; s < params.NUM_SAMPLES; ++s)
                        v[s] = (float) ((double) (int32_t) ((uint32_t) it.first * 2654435761u + (uint32_t) snp * 40503u + s * 2246822519u) / 4294967296.0);
                }
        });
        PHASE("write_decrypted_predictions", write_decrypted_predictions(dec, params, out + "/" RESULT_BYPOS_FILE, false));
    }
    return 0;
}
