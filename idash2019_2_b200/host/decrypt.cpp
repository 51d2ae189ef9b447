// decrypt -- the `decrypt [bypos]` stage (eval/decrypt.cpp:13-49) on the B200 host layer: keys.bin +
// encrypted_prediction.bin -> result.csv (target names) or, with any argument, result_bypos.csv (target positions).
#include "idash_host.h"

int main(int argc, char **) {
    Profiler profiler;
    IdashKey key;
    EncryptedPredictions enc_predictions;
    DecryptedPredictions dec_predictions;

    read_key(key, KEYS_FILE);
    read_encrypted_predictions(enc_predictions, *key.idashParams, ENCRYPTED_PREDICTION_FILE);
    const double t0 = profiler.walltime();
    decrypt_predictions(dec_predictions, enc_predictions, key);
    const double t1 = profiler.walltime();
    if (argc == 1) write_decrypted_predictions(dec_predictions, *key.idashParams, RESULT_FILE, true);
    else write_decrypted_predictions(dec_predictions, *key.idashParams, RESULT_BYPOS_FILE, false);
    const double t_end = profiler.walltime();

    std::cout << "----------------- BENCHMARK ----------------- " << std::endl;
    std::cout << "decrypt wall time (seconds)......: " << t1 - t0 << std::endl;
    std::cout << "serialization wall time (seconds): " << t_end - t1 + t0 << std::endl;
    std::cout << "total wall time (seconds)........: " << t_end << std::endl;
    std::cout << "RAM usage (MB)...................: " << profiler.maxrss() / 1e6 << std::endl;
    std::cout << "gpu call wall time (seconds).....: " << idash_host_last_gpu_seconds() << std::endl;
    return 0;
}
