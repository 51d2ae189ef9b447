// decrypt -- the `decrypt [bypos]` stage (eval/decrypt.cpp:13-49) on the B200 host layer: keys.bin +
// encrypted_prediction.bin -> result.csv (target names) or, with any argument, result_bypos.csv (target positions).
#include "idash_host.h"

int main(int argc, char **) {
    const StageClock clock;
    const bool by_position = argc > 1;
    IdashKey key;
    EncryptedPredictions ciphertexts;
    DecryptedPredictions scores;

    read_key(key, KEYS_FILE);
    const IdashParams &params = *key.idashParams;
    read_encrypted_predictions(ciphertexts, params, ENCRYPTED_PREDICTION_FILE);
    const double stage_s = clock.time([&] { decrypt_predictions(scores, ciphertexts, key); });
    write_decrypted_predictions(scores, params, by_position ? RESULT_BYPOS_FILE : RESULT_FILE, !by_position);
    clock.print_benchmark("decrypt wall time (seconds)......: ", stage_s);
    return 0;
}
