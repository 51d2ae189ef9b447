// idash_host.h -- host layer of the B200 evaluator: the reference's L2 API for the cloud / decrypt stages
// (eval/idash.h) re-implemented over libidash_b200.so, so that `cloud` and `decrypt` mains written against the
// reference header compile and behave the same: same type and member names, same argument meaning, same
// on-disk formats, same error convention ("ERROR: ..." on stdout + abort(), eval/idash.h:12-14).
//
// What is different underneath (and why this is not the reference's header):
//   * a ciphertext file is ONE host slab (`CtSlab`) holding the record stream exactly as it is on disk
//     (eval/idash.cpp:540-556, 596-613); the TLweSample objects in the hash maps are views into that slab, so
//     the C ABI consumes / produces the file image directly (IDASH_B200_LAYOUT_RECORDS) with no
//     per-polynomial allocation or copy. Containers built by other code (separately allocated samples) still
//     work: they are gathered into a packed staging buffer first.
//   * cloud_compute_score / decrypt_predictions have no arithmetic in them: they flatten the Model to CSR,
//     call idash_b200_cloud_eval_host / idash_b200_decrypt_host and fan the results back into the containers.
//   * no TFHE library is needed: tfhe_min.h carries the five plain structs with the reference's member names. (TFHE's own
//     TorusPolynomial / TLweSample own their arrays and have const members, so they cannot be views into a slab; code that must keep
//     <tfhe.h>'s types binds the C ABI through host/reference_glue/idash_b200_glue.cpp instead -- INTEGRATION.md section B.)
#ifndef IDASH_B200_HOST_H
#define IDASH_B200_HOST_H

#include <array>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "tfhe_min.h"

#define REQUIRE_DRAMATICALLY(cond, message) \
    do { if (!(cond)) { std::cout << "ERROR: " << message << std::endl; abort(); } } while (0)
#define DIE_DRAMATICALLY(message) \
    do { std::cout << "ERROR: " << message << std::endl; abort(); } while (0)

// default file names of the pipeline stages (eval/idash.h:16-34)
#define MODEL_FILE "../../ml/model/final"
#define PARAMS_FILE "params.bin"
#define KEYS_FILE "keys.bin"
#define ENCRYPTED_DATA_FILE "encrypted_data.bin"
#define ENCRYPTED_PREDICTION_FILE "encrypted_prediction.bin"
#define RESULT_FILE "result.csv"
#define RESULT_BYPOS_FILE "result_bypos.csv"

typedef uint32_t FeatBigIndex;   // 3 * line number + variant (eval/idash.cpp:337-339, 388-390)
typedef uint32_t FeatIndex;      // ciphertext index = bigIndex / NUM_REGIONS
typedef uint32_t FeatRegion;     // bigIndex % NUM_REGIONS

// eval/idash.h:45-111. Only what the cloud / decrypt stages read.
struct IdashParams {
    static const uint32_t NUM_SNP_PER_POSITIONS = 3;
    static const uint32_t N = 1024;
    static const uint32_t k = 1;
    static const double alpha;               // 2^-25: standard deviation of fresh ciphertexts (eval/idash.cpp:20-21)
    static const TLweParams *tlweParams;
    static const Torus32 ONE_IN_T32;         // 262144 (eval/idash.cpp:29-39)

    uint32_t NUM_SAMPLES = 0;
    uint32_t NUM_INPUT_POSITIONS = 0;
    uint32_t NUM_OUTPUT_POSITIONS = 0;
    uint32_t NUM_INPUT_FEATURES = 0;
    uint32_t NUM_OUTPUT_FEATURES = 0;
    uint32_t NUM_REGIONS = 0;
    uint32_t REGION_SIZE = 0;

    std::unordered_map<uint64_t, std::array<FeatBigIndex, 3>> in_features_index;
    std::unordered_map<uint64_t, std::array<FeatBigIndex, 3>> out_features_index;
    std::vector<std::pair<uint64_t, std::string>> out_position_names;   // order of the result csv

    FeatIndex feature_indexOf(uint32_t big_index) const { return big_index / NUM_REGIONS; }
    FeatRegion feature_regionOf(uint32_t big_index) const { return big_index % NUM_REGIONS; }
    FeatBigIndex feature_bigIndexOf(uint32_t index, uint32_t region) const { return index * NUM_REGIONS + region; }
    FeatBigIndex constant_bigIndex() const { return 0xFFFFFFFFu; }
    // std::out_of_range on an unknown position, like the reference's .at() chain
    FeatBigIndex inBigIdx(const uint64_t &pos, uint64_t snp) const { return in_features_index.at(pos).at(snp); }
    FeatBigIndex outBigIdx(const uint64_t &pos, uint64_t snp) const { return out_features_index.at(pos).at(snp); }
};

struct IdashKey {
    const IdashParams *idashParams = nullptr;
    const TLweKey *tlweKey = nullptr;
    IdashKey() {}
    IdashKey(const IdashParams *p, const TLweKey *k) : idashParams(p), tlweKey(k) {}
};

// model[output bigIndex][input bigIndex or constant_bigIndex()] = integer coefficient (eval/idash.h:129-134)
struct CompiledModel;   // idash_host.cpp: the block-banded device layout read_model emits
struct Model {
    std::unordered_map<FeatBigIndex, std::unordered_map<FeatBigIndex, int32_t>> model;
    // B200: set by read_model -- the model compiled into the device layout (and uploaded), with the order in which the reference
    // would walk `model`. When present it is what cloud_compute_score evaluates; `model` itself is left EMPTY when read_model took
    // the layout from the model cache (models.bin) instead of parsing the .hr files. Hand-built / edited models: compiled.reset().
    std::shared_ptr<CompiledModel> compiled;
};

// One host allocation holding `count` ciphertext records in file layout, preceded by the u64 count:
//   image() = { u64 count, count x { u32 index, i32 84, u32 a[1024], u32 b[1024], f64 variance } }
// image() is 16-byte aligned, hence records() is 8 mod 16 and every word array is 16-byte aligned.
struct CtSlab {
    uint64_t count = 0;
    uint8_t *mem = nullptr;                 // pre-faulted huge-page mapping (pageable); see host_buf_alloc in idash_host.cpp
    int mem_kind = 0;
    size_t mem_bytes = 0;
    bool registered = false;                // page-locked through idash_b200_host_register (unregistered by the destructor)
    std::vector<uint32_t> sorted_index;     // non-empty: the records are sorted by ciphertext index and this is index_of(i) for every slot
    std::vector<TLweSample> samples;        // views: samples[i].a[0].coefsT / b->coefsT point into record i
    std::vector<TorusPolynomial> polys;

    explicit CtSlab(uint64_t count);
    ~CtSlab();
    CtSlab(const CtSlab &) = delete;
    CtSlab &operator=(const CtSlab &) = delete;
    uint8_t *image() const { return mem; }
    uint8_t *records() const { return mem + 8; }
    size_t image_bytes() const { return 8 + (size_t) count * 8208u; }
    uint8_t *record(uint64_t i) const { return mem + 8 + i * 8208u; }
    uint32_t index_of(uint64_t i) const;
    // record i -> samples[i].current_variance (after the device / a file read filled the records)
    void pull_variances();
    // samples[i].current_variance -> record i (before writing the image out)
    void push_variances();
    // slot of a sample that is a view into this slab, or -1
    int64_t slot_of(const TLweSample *s) const {
        return (s >= samples.data() && s < samples.data() + count) ? (int64_t) (s - samples.data()) : -1;
    }
};

struct EncryptedData {
    std::unordered_map<FeatIndex, TLweSample *> enc_data;
    std::shared_ptr<CtSlab> slab;           // set by read_encrypted_data; may be empty for hand-built containers

    const TLweSample *getTLWE(FeatBigIndex inBidx, const IdashParams &params) const {
        const auto it = enc_data.find(params.feature_indexOf(inBidx));
        REQUIRE_DRAMATICALLY(it != enc_data.end(), "shit happens before");   // message of eval/idash.h:164
        return it->second;
    }
};

struct EncryptedPredictions {
    std::unordered_map<FeatBigIndex, TLweSample *> score;
    std::shared_ptr<CtSlab> slab;           // set by cloud_compute_score / read_encrypted_predictions

    TLweSample *get(FeatBigIndex bidx, const TLweParams *) { return score.at(bidx); }
};

struct DecryptedPredictions {
    // score[target position][variant][sample]
    std::unordered_map<uint64_t, std::array<std::vector<float>, 3>> score;
};

void read_params(IdashParams &params, const std::string &filename);
void read_key(IdashKey &key, const std::string &filename);
void read_model(Model &model, const IdashParams &params, const std::string &path);
void read_encrypted_data(EncryptedData &encrypted_data, const IdashParams &params, const std::string &filename);
void write_encrypted_data(const EncryptedData &encrypted_data, const IdashParams &params, const std::string &filename);
void read_encrypted_predictions(EncryptedPredictions &encrypted_preds, const IdashParams &params, const std::string &filename);
void write_encrypted_predictions(const EncryptedPredictions &encrypted_preds, const IdashParams &params, const std::string &filename);
void write_decrypted_predictions(const DecryptedPredictions &predictions, const IdashParams &params, const std::string &filename,
                                 const bool PRINT_POS_NAME = true);

// eval/idash.cpp:763-848, on the GPU. enc_preds.score must be empty; on return it holds one ciphertext per model
// row, inserted in the reference's order, so write_encrypted_predictions produces the reference's file.
void cloud_compute_score(EncryptedPredictions &enc_preds, const EncryptedData &enc_data, const Model &model, const IdashParams &params);
// eval/idash.cpp:681-761, on the GPU, with the exact integer phase (the reference's FFT is within 1 LSB of it).
void decrypt_predictions(DecryptedPredictions &predictions, const EncryptedPredictions &enc_preds, const IdashKey &key);

// GPU used by this process: IDASH_B200_DEVICE, else LOCAL_RANK, else 0.
int idash_host_device();
// GPUs cloud_compute_score shards the target range over: IDASH_GPUS="0,1,2,3" (default: the one of idash_host_device()).
std::vector<int> idash_host_devices();
// The cached packed model: read_model stores / finds the compiled layout in this file ("" = off). The `cloud` binary
// turns it on with "models.bin" in the working directory unless IDASH_MODEL_CACHE is set ("0" / "off" disables, else a path).
void idash_host_set_model_cache(const std::string &path);
// Blocks until the helper threads read_params / read_model / read_encrypted_data started (CUDA contexts, output slab, page-locking,
// model upload) are done, so that a stage timer started afterwards measures the evaluation and not process start-up.
void idash_host_wait_ready();
// seconds spent inside the C-ABI call of the last cloud_compute_score / decrypt_predictions (for the BENCHMARK block)
double idash_host_last_gpu_seconds();

// eval/idash.h:272-292
class Profiler {
public:
    Profiler();
    static double universalWallTime();
    static double universalClockTime();
    double walltime() const;
    double clocktime() const;
    long int maxrss() const;   // bytes
private:
    const double tw0, tc0;
};


// Timing of a stage binary in the reference's BENCHMARK format (eval/cloud.cpp:21-26, eval/decrypt.cpp:42-47,
// parsed by eval/parse_log.py): the stage itself, everything else as "serialization", the total, the peak RSS --
// plus the time spent inside the C-ABI calls.
class StageClock {
public:
    template <class F>
    double time(F &&stage) const {
        const double t0 = profiler.walltime();
        stage();
        return profiler.walltime() - t0;
    }
    void print_benchmark(const char *stage_label, double stage_seconds) const;
private:
    Profiler profiler;
};

#endif
