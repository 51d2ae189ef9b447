// cloud -- the `cloud <model_dir>` stage (eval/cloud.cpp:3-28) on the B200 host layer: same inputs in the working
// directory (params.bin, encrypted_data.bin), same output (encrypted_prediction.bin), same BENCHMARK block on stdout
// (eval/parse_log.py:14-41 keeps working), plus one line with the time spent inside the GPU calls.
#include "idash_host.h"

namespace {
struct CloudRun {
    IdashParams params;
    Model model;
    EncryptedData inputs;
    EncryptedPredictions outputs;
};
}  // namespace

int main(int argc, char **argv) {
    const StageClock clock;
    const std::string model_dir = argc > 1 ? argv[1] : MODEL_FILE;
    std::cout << "using model dir: " << model_dir << std::endl;

    // the cached packed model: "models.bin" next to the other files of the stage unless IDASH_MODEL_CACHE says otherwise
    const char *cache = getenv("IDASH_MODEL_CACHE");
    if (!cache) idash_host_set_model_cache("models.bin");
    else if (std::string(cache) != "0" && std::string(cache) != "off" && *cache) idash_host_set_model_cache(cache);

    CloudRun run;
    read_params(run.params, PARAMS_FILE);          // also starts the CUDA context(s) on a helper thread
    read_model(run.model, run.params, model_dir);  // ... the output slab, and the upload of the compiled model
    read_encrypted_data(run.inputs, run.params, ENCRYPTED_DATA_FILE);
    // process start-up (CUDA context creation, page-locking) is not the evaluation: like the reference's timer, `fhe wall time` is
    // cloud_compute_score alone; everything else -- including this wait -- lands in the `serialization` and `total` lines
    idash_host_wait_ready();
    const double stage_s = clock.time([&] { cloud_compute_score(run.outputs, run.inputs, run.model, run.params); });
    write_encrypted_predictions(run.outputs, run.params, ENCRYPTED_PREDICTION_FILE);
    clock.print_benchmark("fhe wall time (seconds)..........: ", stage_s);
    return 0;
}
