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

    CloudRun run;
    read_params(run.params, PARAMS_FILE);          // also starts the CUDA context on a helper thread
    read_model(run.model, run.params, model_dir);  // ... and the allocation of the output slab
    read_encrypted_data(run.inputs, run.params, ENCRYPTED_DATA_FILE);
    const double stage_s = clock.time([&] { cloud_compute_score(run.outputs, run.inputs, run.model, run.params); });
    write_encrypted_predictions(run.outputs, run.params, ENCRYPTED_PREDICTION_FILE);
    clock.print_benchmark("fhe wall time (seconds)..........: ", stage_s);
    return 0;
}
