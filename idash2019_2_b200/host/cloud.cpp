// cloud -- the `cloud <model_dir>` stage (eval/cloud.cpp:3-28) on the B200 host layer: same inputs in the working
// directory (params.bin, encrypted_data.bin), same output (encrypted_prediction.bin), same BENCHMARK block on stdout
// (eval/parse_log.py:14-41 keeps working), plus one line with the time spent inside the GPU call.
#include "idash_host.h"

int main(int argc, char **argv) {
    Profiler profiler;
    std::string modelFile = MODEL_FILE;
    if (argc >= 2) modelFile = argv[1];
    std::cout << "using model dir: " << modelFile << std::endl;

    IdashParams params;
    Model model;
    EncryptedData enc_data;
    EncryptedPredictions enc_preds;
    read_params(params, PARAMS_FILE);
    read_model(model, params, modelFile);
    read_encrypted_data(enc_data, params, ENCRYPTED_DATA_FILE);
    const double t_cloud0 = profiler.walltime();
    cloud_compute_score(enc_preds, enc_data, model, params);
    const double t_cloud1 = profiler.walltime();
    write_encrypted_predictions(enc_preds, params, ENCRYPTED_PREDICTION_FILE);
    const double t_end = profiler.walltime();

    std::cout << "----------------- BENCHMARK ----------------- " << std::endl;
    std::cout << "fhe wall time (seconds)..........: " << t_cloud1 - t_cloud0 << std::endl;
    std::cout << "serialization wall time (seconds): " << t_end - t_cloud1 + t_cloud0 << std::endl;
    std::cout << "total wall time (seconds)........: " << t_end << std::endl;
    std::cout << "RAM usage (MB)...................: " << profiler.maxrss() / 1e6 << std::endl;
    std::cout << "gpu call wall time (seconds).....: " << idash_host_last_gpu_seconds() << std::endl;
    return 0;
}
