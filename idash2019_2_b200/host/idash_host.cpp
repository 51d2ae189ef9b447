// idash_host.cpp -- see idash_host.h. Host logic only: file formats, container plumbing and the calls into
// libidash_b200.so. All torus arithmetic of the cloud and decrypt stages runs on the GPU behind the C ABI; there
// is no CPU evaluation path in this file.
#include "idash_host.h"
#include "parse_vw.h"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <future>
#include <mutex>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <sys/resource.h>
#include <sys/time.h>
#include <time.h>

#include "idash_b200.h"

using std::string;

// ---- constants (eval/idash.cpp:20-45) --------------------------------------------------------------------------
const double IdashParams::alpha = 1.0 / 33554432.0;            // pow(2., -25)
const Torus32 IdashParams::ONE_IN_T32 = IDASH_B200_ONE_IN_T32;   // dtot32(1 / 16384)
static const TLweParams g_tlwe_params = {(int32_t) IdashParams::N, (int32_t) IdashParams::k, IdashParams::alpha, 0.25};
const TLweParams *IdashParams::tlweParams = &g_tlwe_params;

static const uint32_t REC = IDASH_B200_RECORD_BYTES;
static const uint32_t REC_WORDS_OFF = 8;                          // u32 index, i32 type uid
static const uint32_t REC_VAR_OFF = 8 + IDASH_B200_CT_BYTES;

namespace {

double g_last_gpu_seconds = 0.0;

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

unsigned host_threads() {
    if (const char *e = getenv("IDASH_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return (unsigned) v; }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 4u;
}

// runs fn(begin, end) over [0, n) split into contiguous ranges, one per worker thread
template <class F>
void parallel_ranges(size_t n, F fn, size_t grain = 64) {
    const size_t nt = std::max<size_t>(1, std::min<size_t>(host_threads(), n / grain + (grain > 1)));
    if (nt == 1) { fn((size_t) 0, n); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

// Host staging memory for ciphertext slabs and score matrices. Default: anonymous mapping with transparent huge pages,
// pre-faulted by all threads -- pageable. Measured on the B200 box for the 2 GB output slab (tools/pin_bench.cu): cudaHostAlloc
// 0.93 s (+ 0.035 s D2H), mmap + touch + cudaHostRegister 0.20 s (+ 0.036 s), mmap + touch and a pageable D2H 0.045 + 0.105 s.
// For a buffer that is written or read by the device ONCE, page-locking costs more than it saves, and cudaHostAlloc holds a
// driver lock that stalls the concurrent model upload. IDASH_HOST_PIN=alloc restores cudaHostAlloc.
enum { HOST_MEM_MALLOC = 0, HOST_MEM_PINNED = 1, HOST_MEM_MAPPED = 2 };
struct HostBuf { void *p; int kind; size_t bytes; };
HostBuf host_buf_alloc(size_t bytes) {
    void *p = nullptr;
    const char *pin = getenv("IDASH_HOST_PIN");
    if (pin && !strcmp(pin, "alloc") && idash_b200_host_alloc(&p, bytes) == IDASH_B200_OK) return {p, HOST_MEM_PINNED, bytes};
    if (bytes >= ((size_t) 4 << 20)) {
        const size_t huge = (size_t) 2 << 20, len = (bytes + huge - 1) & ~(huge - 1);
        p = ::mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            ::madvise(p, len, MADV_HUGEPAGE);
            uint8_t *b = static_cast<uint8_t *>(p);
            parallel_ranges(len / huge, [&](size_t lo, size_t hi) {
                for (size_t i = lo * huge; i < hi * huge; i += 4096) b[i] = 0;     // fault the pages in, all threads at once
            }, 1);
            return {p, HOST_MEM_MAPPED, len};
        }
    }
    p = aligned_alloc(256, (bytes + 255) & ~(size_t) 255);
    REQUIRE_DRAMATICALLY(p != nullptr, "out of host memory (" << bytes << " bytes)");
    return {p, HOST_MEM_MALLOC, bytes};
}
void host_buf_free(const HostBuf &b) {
    if (!b.p) return;
    if (b.kind == HOST_MEM_PINNED) idash_b200_host_free(b.p);
    else if (b.kind == HOST_MEM_MAPPED) ::munmap(b.p, b.bytes);
    else free(b.p);
}

// one context per process (one process per GPU)
// (thread-safe: the cloud binary creates it on a helper thread while the model files are parsed; `die` = false is that
// helper's probe -- a box without a GPU must still run the file-format half of this layer)
idash_b200_ctx *gpu_ctx(bool die = true) {
    static idash_b200_ctx *ctx = nullptr;
    static string error;
    static std::once_flag once;
    std::call_once(once, []() {
        if (idash_b200_init(&ctx, idash_host_device()) != IDASH_B200_OK) { ctx = nullptr; error = idash_b200_last_error(); }
    });
    if (!ctx && die) DIE_DRAMATICALLY("idash_b200_init: " << error);
    return ctx;
}

// Whole-range pread / pwrite by a pool of threads: a 2 GB ciphertext file moves through the page cache at memory speed
// instead of one core's memcpy speed (eval/idash.cpp:513-613 streams record by record).
void parallel_file_io(int fd, uint8_t *buf, size_t bytes, off_t off0, bool write, const char *what) {
    std::atomic<bool> bad(false);
    const size_t piece = (size_t) 8 << 20;
    const size_t n_pieces = (bytes + piece - 1) / piece;
    std::atomic<size_t> next(0);
    const size_t nt = std::max<size_t>(1, std::min<size_t>(host_threads(), n_pieces));
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n_pieces || bad) return;
                size_t lo = i * piece;
                const size_t hi = std::min(bytes, lo + piece);
                while (lo < hi) {
                    const ssize_t r = write ? ::pwrite(fd, buf + lo, hi - lo, off0 + (off_t) lo) : ::pread(fd, buf + lo, hi - lo, off0 + (off_t) lo);
                    if (r < 0 && errno == EINTR) continue;
                    if (r <= 0) { bad = true; return; }
                    lo += (size_t) r;
                }
            }
        });
    for (auto &x : th) x.join();
    REQUIRE_DRAMATICALLY(!bad, (write ? "short write to encrypted " : "truncated encrypted ") << what << " file");
}

void read_exact(std::istream &in, void *dst, size_t n, const char *what) {
    in.read(static_cast<char *>(dst), (std::streamsize) n);
    REQUIRE_DRAMATICALLY((size_t) in.gcount() == n, "truncated file while reading " << what);
}

// params stream: eval/idash.cpp:128-186 (writer), 200-261 (reader)
void read_params_stream(IdashParams &p, std::istream &in) {
    uint32_t head[7];
    read_exact(in, head, sizeof(head), "params header");
    p.NUM_SAMPLES = head[0]; p.NUM_INPUT_POSITIONS = head[1]; p.NUM_OUTPUT_POSITIONS = head[2];
    p.NUM_INPUT_FEATURES = head[3]; p.NUM_OUTPUT_FEATURES = head[4]; p.NUM_REGIONS = head[5]; p.REGION_SIZE = head[6];
    REQUIRE_DRAMATICALLY(p.NUM_REGIONS >= 1 && p.REGION_SIZE >= 1 && (uint64_t) p.NUM_REGIONS * p.REGION_SIZE <= IdashParams::N,
                         "params: NUM_REGIONS x REGION_SIZE does not fit the polynomial");
    std::vector<uint8_t> tags((size_t) p.NUM_INPUT_POSITIONS * 20u);     // packed {u64 pos, u32 bidx[3]}
    if (!tags.empty()) read_exact(in, tags.data(), tags.size(), "tag positions");
    for (uint32_t i = 0; i < p.NUM_INPUT_POSITIONS; ++i) {
        uint64_t pos;
        std::array<FeatBigIndex, 3> b;
        memcpy(&pos, &tags[(size_t) i * 20u], 8);
        memcpy(b.data(), &tags[(size_t) i * 20u + 8], 12);
        p.in_features_index.emplace(pos, b);           // first occurrence wins, as with the reference's emplace
    }
    REQUIRE_DRAMATICALLY(p.NUM_INPUT_POSITIONS == p.in_features_index.size(), "NUM_INPUT_POSITIONS != in_features_index.size()");
    for (uint32_t i = 0; i < p.NUM_OUTPUT_POSITIONS; ++i) {
        uint64_t pos;
        read_exact(in, &pos, 8, "target position");
        string name;
        std::getline(in, name, '\0');
        REQUIRE_DRAMATICALLY(!in.fail(), "truncated file while reading target name");
        std::array<FeatBigIndex, 3> b;
        read_exact(in, b.data(), 12, "target indices");
        p.out_position_names.push_back({pos, name});
        p.out_features_index.emplace(pos, b);
    }
    REQUIRE_DRAMATICALLY(p.NUM_OUTPUT_POSITIONS == p.out_features_index.size(), "NUM_OUTPUT_POSITIONS != out_features_index.size()");
    REQUIRE_DRAMATICALLY(p.NUM_OUTPUT_POSITIONS == p.out_position_names.size(), "NUM_OUTPUT_POSITIONS != out_position_names.size()");
}

// whole ciphertext file -> slab; checks the size and every record's TLWE type uid (tfhe_io.cpp:308 aborts on a bad one)
// Buffered write()s to ONE file are serialised by the inode lock (measured: 8 threads of pwrite are no faster than one),
// so a large image is copied through a shared mapping instead: page-cache pages are faulted in and filled by all
// threads at once. Falls back to pwrite where the file cannot be mapped.
void parallel_file_write(int fd, uint8_t *buf, size_t bytes, const char *what) {
    if (bytes == 0) return;
    void *map = getenv("IDASH_HOST_NO_MMAP") ? MAP_FAILED : ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (map == MAP_FAILED) { parallel_file_io(fd, buf, bytes, 0, true, what); return; }
    const size_t piece = (size_t) 4 << 20;
    parallel_ranges((bytes + piece - 1) / piece, [&](size_t b, size_t e) {
        const size_t lo = b * piece, hi = std::min(bytes, e * piece);
        if (lo < hi) memcpy(static_cast<uint8_t *>(map) + lo, buf + lo, hi - lo);
    }, 1);
    REQUIRE_DRAMATICALLY(::munmap(map, bytes) == 0, "error unmapping encrypted " << what << " file");
}

std::shared_ptr<CtSlab> read_ct_file(const string &filename, const char *what) {
    const int fd = ::open(filename.c_str(), O_RDONLY);
    REQUIRE_DRAMATICALLY(fd >= 0, "Cannot open encrypted " << what << " file for read");
    uint64_t count = 0;
    REQUIRE_DRAMATICALLY(::pread(fd, &count, 8, 0) == 8, "truncated encrypted " << what << " file");
    REQUIRE_DRAMATICALLY(count < ((uint64_t) 1 << 32), "encrypted " << what << " file: implausible record count");
    auto slab = std::make_shared<CtSlab>(count);
    parallel_file_io(fd, slab->records(), (size_t) count * REC, 8, false, what);
    ::close(fd);
    std::atomic<uint64_t> bad_rec(UINT64_MAX);
    parallel_ranges(count, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            int32_t uid;
            memcpy(&uid, slab->record(i) + 4, 4);
            if (uid != IDASH_B200_TLWE_SAMPLE_UID) { bad_rec = i; return; }
            memcpy(&slab->samples[i].current_variance, slab->record(i) + REC_VAR_OFF, 8);
        }
    });
    REQUIRE_DRAMATICALLY(bad_rec == UINT64_MAX, "encrypted " << what << " file: bad TLWE sample type in record " << bad_rec.load());
    return slab;
}

// map iteration order == slab slot order? then the slab image IS the file
template <class Map>
bool map_matches_slab(const Map &m, const std::shared_ptr<CtSlab> &slab) {
    if (!slab || slab->count != m.size()) return false;
    uint64_t k = 0;
    for (const auto &it : m) {
        if (it.second != &slab->samples[k] || slab->index_of(k) != it.first) return false;
        ++k;
    }
    return true;
}

template <class Map>
void write_ct_file(const Map &m, const std::shared_ptr<CtSlab> &slab, const string &filename, const char *what) {
    if (map_matches_slab(m, slab)) {   // the slab image IS the file: one parallel write
        slab->push_variances();
        const int fd = ::open(filename.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        REQUIRE_DRAMATICALLY(fd >= 0, "Cannot open encrypted " << what << " file for write");
        REQUIRE_DRAMATICALLY(::ftruncate(fd, (off_t) slab->image_bytes()) == 0, "cannot size encrypted " << what << " file");
        parallel_file_write(fd, slab->image(), slab->image_bytes(), what);
        REQUIRE_DRAMATICALLY(::close(fd) == 0, "error closing encrypted " << what << " file");
        return;
    }
    FILE *f = fopen(filename.c_str(), "wb");
    REQUIRE_DRAMATICALLY(f != nullptr, "Cannot open encrypted " << what << " file for write");
    {   // containers not built by this library: record by record, in the map's iteration order
        const uint64_t count = m.size();
        fwrite(&count, 8, 1, f);
        std::vector<uint8_t> rec(REC);
        const int32_t uid = IDASH_B200_TLWE_SAMPLE_UID;
        for (const auto &it : m) {
            const uint32_t idx = it.first;
            memcpy(&rec[0], &idx, 4);
            memcpy(&rec[4], &uid, 4);
            memcpy(&rec[REC_WORDS_OFF], it.second->a[0].coefsT, 4096);
            memcpy(&rec[REC_WORDS_OFF + 4096], it.second->a[1].coefsT, 4096);
            memcpy(&rec[REC_VAR_OFF], &it.second->current_variance, 8);
            REQUIRE_DRAMATICALLY(fwrite(rec.data(), 1, REC, f) == REC, "short write to encrypted " << what << " file");
        }
    }
    REQUIRE_DRAMATICALLY(fclose(f) == 0, "error closing encrypted " << what << " file");
}

// pinned staging buffer for containers whose samples are not views into a slab
struct PackedCts {
    void *mem = nullptr;
    HostBuf buf{nullptr, 0, 0};
    uint64_t count = 0;
    uint32_t *words() const { return static_cast<uint32_t *>(mem); }
    uint32_t *index() const { return words() + count * IDASH_B200_CT_WORDS; }
    double *variance() const { return reinterpret_cast<double *>(static_cast<uint8_t *>(mem) + ((count * (IDASH_B200_CT_BYTES + 4u) + 7u) & ~(uint64_t) 7u)); }
    explicit PackedCts(uint64_t n) : count(n) {
        const size_t bytes = ((n * (IDASH_B200_CT_BYTES + 4u) + 7u) & ~(uint64_t) 7u) + n * 8u + 16u;
        buf = host_buf_alloc(bytes);
        mem = buf.p;
    }
    ~PackedCts() { host_buf_free(buf); }
};

template <class Map>
std::unique_ptr<PackedCts> pack_map(const Map &m) {
    std::unique_ptr<PackedCts> pk(new PackedCts(m.size()));
    std::vector<std::pair<uint32_t, const TLweSample *>> v;
    v.reserve(m.size());
    for (const auto &it : m) v.push_back({it.first, it.second});
    parallel_ranges(v.size(), [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            memcpy(pk->words() + i * IDASH_B200_CT_WORDS, v[i].second->a[0].coefsT, 4096);
            memcpy(pk->words() + i * IDASH_B200_CT_WORDS + 1024, v[i].second->a[1].coefsT, 4096);
            pk->index()[i] = v[i].first;
            pk->variance()[i] = v[i].second->current_variance;
        }
    });
    return pk;
}

// Helper-thread warm-up for the cloud stage (started by read_model, collected by cloud_compute_score): GPU context +
// the slab the output ciphertexts will be copied into.
std::mutex g_warm_mutex;
std::shared_future<void> g_warm_ctx;
std::future<std::shared_ptr<CtSlab>> g_warm_slab;
// read_params is the first call of both GPU stages: start creating the CUDA context (0.3-0.5 s) right away
void warm_up_context() {
    if (getenv("IDASH_HOST_NO_WARMUP")) return;
    std::lock_guard<std::mutex> lock(g_warm_mutex);
    if (g_warm_ctx.valid()) return;
    g_warm_ctx = std::async(std::launch::async, []() {
        const double t0 = now_s();
        const bool ok = gpu_ctx(false) != nullptr;
        if (getenv("IDASH_HOST_TIMING")) fprintf(stderr, "[idash_host] warm-up: context %s in %.3f s\n", ok ? "created" : "unavailable", now_s() - t0);
    }).share();
}
void warm_up_cloud(uint64_t n_rows) {
    if (getenv("IDASH_HOST_NO_WARMUP")) return;
    warm_up_context();
    std::lock_guard<std::mutex> lock(g_warm_mutex);
    if (g_warm_slab.valid()) return;
    g_warm_slab = std::async(std::launch::async, [n_rows]() {
        const double t0 = now_s();
        auto slab = std::make_shared<CtSlab>(n_rows);
        if (getenv("IDASH_HOST_TIMING")) fprintf(stderr, "[idash_host] warm-up: output slab (%.2f GB) in %.3f s\n", slab->image_bytes() * 1e-9, now_s() - t0);
        return slab;
    });
}
std::shared_ptr<CtSlab> take_output_slab(uint64_t n_rows) {
    std::shared_ptr<CtSlab> slab;
    {
        std::lock_guard<std::mutex> lock(g_warm_mutex);
        if (g_warm_slab.valid()) slab = g_warm_slab.get();
    }
    if (!slab || slab->count != n_rows) slab = std::make_shared<CtSlab>(n_rows);
    return slab;
}

template <class Map>
bool all_views_of(const Map &m, const std::shared_ptr<CtSlab> &slab) {
    if (!slab || slab->count != m.size()) return false;
    for (const auto &it : m) {
        const int64_t s = slab->slot_of(it.second);
        if (s < 0 || slab->index_of((uint64_t) s) != it.first) return false;
    }
    return true;
}

}  // namespace

// ---- CtSlab ----------------------------------------------------------------------------------------------------
CtSlab::CtSlab(uint64_t n) : count(n), samples(n), polys(2 * n) {
    const HostBuf hb = host_buf_alloc(image_bytes() + 16);
    mem = static_cast<uint8_t *>(hb.p);   // at least 256-byte aligned either way
    mem_kind = hb.kind;
    mem_bytes = hb.bytes;
    memcpy(mem, &count, 8);
    for (uint64_t i = 0; i < n; ++i) {
        polys[2 * i].N = polys[2 * i + 1].N = (int32_t) IdashParams::N;
        polys[2 * i].coefsT = reinterpret_cast<Torus32 *>(record(i) + REC_WORDS_OFF);
        polys[2 * i + 1].coefsT = polys[2 * i].coefsT + IdashParams::N;
        samples[i].a = &polys[2 * i];
        samples[i].b = &polys[2 * i + 1];
        samples[i].current_variance = 0.0;
        samples[i].k = (int32_t) IdashParams::k;
    }
}
CtSlab::~CtSlab() { host_buf_free(HostBuf{mem, mem_kind, mem_bytes}); }
uint32_t CtSlab::index_of(uint64_t i) const { uint32_t v; memcpy(&v, record(i), 4); return v; }
void CtSlab::pull_variances() { for (uint64_t i = 0; i < count; ++i) memcpy(&samples[i].current_variance, record(i) + REC_VAR_OFF, 8); }
void CtSlab::push_variances() { for (uint64_t i = 0; i < count; ++i) memcpy(record(i) + REC_VAR_OFF, &samples[i].current_variance, 8); }

// ---- files -----------------------------------------------------------------------------------------------------
// NOTE: none of the reference-visible hash maps is reserve()d anywhere in this file: their bucket counts, hence their
// iteration order, must evolve exactly as in the reference, because that order is the record order of the files it
// writes (eval/idash.cpp:550, 607) and the row order of the model walk (eval/idash.cpp:772).
void read_params(IdashParams &params, const string &filename) {
    warm_up_context();
    std::ifstream in(filename.c_str(), std::ios::binary);
    REQUIRE_DRAMATICALLY(in.is_open(), "Cannot open parameters file for read");
    read_params_stream(params, in);
}

// keys.bin = params stream, TLWEPARAMS text block, i32 85, N key bits (eval/idash.cpp:476-502, tfhe_io.cpp:243-252, 395-436)
void read_key(IdashKey &key, const string &filename) {
    std::ifstream in(filename.c_str(), std::ios::binary);
    REQUIRE_DRAMATICALLY(in.is_open(), "Cannot open key file for read");
    IdashParams *params = new IdashParams();
    read_params_stream(*params, in);
    key.idashParams = params;
    int64_t N = -1, k = -1;
    double a_min = 0, a_max = 0;
    string line;
    REQUIRE_DRAMATICALLY(std::getline(in, line) && line == "-----BEGIN TLWEPARAMS-----", "key file: TLWEPARAMS block not found");
    for (;;) {
        REQUIRE_DRAMATICALLY((bool) std::getline(in, line), "key file: unterminated TLWEPARAMS block");
        if (line == "-----END TLWEPARAMS-----") break;
        const size_t c = line.find(": ");
        if (c == string::npos) continue;
        const string name = line.substr(0, c), val = line.substr(c + 2);
        if (name == "N") N = atoll(val.c_str());
        else if (name == "k") k = atoll(val.c_str());
        else if (name == "alpha_min") a_min = atof(val.c_str());
        else if (name == "alpha_max") a_max = atof(val.c_str());
    }
    REQUIRE_DRAMATICALLY(N == (int64_t) IdashParams::N && k == 1, "key file: unsupported TLWE parameters N=" << N << " k=" << k);
    int32_t uid = 0;
    read_exact(in, &uid, 4, "key type");
    REQUIRE_DRAMATICALLY(uid == 85, "key file: bad TLWE key type");         // TLWE_KEY_TYPE_UID, tfhe_generic_streams.h:27
    TLweParams *tp = new TLweParams{(int32_t) N, (int32_t) k, a_min, a_max};
    IntPolynomial *poly = new IntPolynomial{(int32_t) N, new int32_t[N]};
    read_exact(in, poly->coefs, sizeof(int32_t) * (size_t) N, "key bits");
    key.tlweKey = new TLweKey{tp, poly};
}

void read_encrypted_data(EncryptedData &d, const IdashParams &, const string &filename) {
    d.slab = read_ct_file(filename, "data");
    for (uint64_t i = 0; i < d.slab->count; ++i) d.enc_data.emplace(d.slab->index_of(i), &d.slab->samples[i]);
}

void read_encrypted_predictions(EncryptedPredictions &p, const IdashParams &, const string &filename) {
    p.slab = read_ct_file(filename, "predictions");
    for (uint64_t i = 0; i < p.slab->count; ++i) p.score.emplace(p.slab->index_of(i), &p.slab->samples[i]);
}

void write_encrypted_data(const EncryptedData &d, const IdashParams &, const string &filename) {
    write_ct_file(d.enc_data, d.slab, filename, "data");
}

void write_encrypted_predictions(const EncryptedPredictions &p, const IdashParams &, const string &filename) {
    write_ct_file(p.score, p.slab, filename, "predictions");
}

// eval/idash.cpp:66-90. The .hr files are parsed by a pool of threads (3 x NUM_OUTPUT_POSITIONS small files);
// the model map is then filled in the reference's order (iteration order of out_features_index, variant 0..2),
// which fixes the record order of encrypted_prediction.bin downstream.
void read_model(Model &model, const IdashParams &params, const string &path) {
    // Only the cloud stage loads a model: bring the GPU context up and allocate + fault in the output slab on helper
    // threads while the files are parsed (CUDA initialisation and 2 GB of fresh host memory are the two largest fixed
    // costs of cloud_compute_score at iDASH scale).
    warm_up_cloud(3 * (uint64_t) params.out_features_index.size());

    const double t_begin = now_s();
    struct Job { uint64_t pos; std::array<FeatBigIndex, 3> out; };
    std::vector<Job> jobs;
    jobs.reserve(params.out_features_index.size());
    for (const auto &e : params.out_features_index) jobs.push_back({e.first, e.second});
    std::vector<std::array<std::unordered_map<FeatBigIndex, int32_t>, 3>> parsed(jobs.size());
    std::atomic<bool> unknown(false);
    parallel_ranges(jobs.size(), [&](size_t b, size_t e) {
        string fn;
        for (size_t i = b; i < e; ++i) {
            for (int snp = 0; snp < 3; ++snp) {
                fn = path; fn += '/'; fn += std::to_string(jobs[i].pos); fn += '_'; fn += (char) ('0' + snp); fn += ".hr";
                auto &dst = parsed[i][snp];
                for (const auto &c : read_lines(fn)) {
                    if (c.first == "Constant") { dst[params.constant_bigIndex()] = c.second; continue; }
                    const size_t u = c.first.find('_');
                    char *end = nullptr;
                    const uint64_t pos = strtoull(c.first.c_str(), &end, 10);
                    const auto it = params.in_features_index.find(pos);
                    const long v = u == string::npos ? -1 : strtol(c.first.c_str() + u + 1, nullptr, 10);
                    if (it == params.in_features_index.end() || v < 0 || v > 2) { unknown = true; continue; }
                    dst[it->second[(size_t) v]] = c.second;                       // a repeated name overwrites, as in parse_vw
                }
            }
        }
    });
    // the reference throws std::out_of_range (uncaught -> abort) on a tag position that params.bin does not know
    REQUIRE_DRAMATICALLY(!unknown, "model refers to a tag SNP feature that is not in the parameters file");
    const double t_parsed = now_s();
    // outer map: filled serially in the reference's order (its iteration order is the row order downstream)
    for (size_t i = 0; i < jobs.size(); ++i)
        for (int snp = 0; snp < 3; ++snp) model.model[jobs[i].out[snp]] = std::move(parsed[i][snp]);
    if (getenv("IDASH_HOST_TIMING")) fprintf(stderr, "[idash_host] read_model: parse %.3f s, fill %.3f s\n", t_parsed - t_begin, now_s() - t_parsed);
}

// eval/idash.cpp:436-469: header, then for every sample, for every target in file order, "<sample>,<target>,<p0>,<p1>,<p2>".
// Floats print like `ostream << float` (%g with 6 significant digits). Formatted by a pool of threads into per-range
// buffers and written with large writes instead of one flush per row.
void write_decrypted_predictions(const DecryptedPredictions &predictions, const IdashParams &params, const string &filename,
                                 const bool PRINT_POS_NAME) {
    const int fd = ::open(filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    REQUIRE_DRAMATICALLY(fd >= 0, "Cannot open result file for write");
    static const char header[] = "Subject ID,target SNP,0,1,2\n";
    off_t file_off = (off_t) (sizeof(header) - 1);
    REQUIRE_DRAMATICALLY(::pwrite(fd, header, sizeof(header) - 1, 0) == (ssize_t) (sizeof(header) - 1), "short write to result file");
    const size_t G = params.out_position_names.size();
    const size_t S = params.NUM_SAMPLES;
    std::vector<const std::array<std::vector<float>, 3> *> rows(G);
    std::vector<string> labels(G);          // ",<label>,"
    size_t label_max = 0;
    for (size_t g = 0; g < G; ++g) {
        rows[g] = &predictions.score.at(params.out_position_names[g].first);
        labels[g] = "," + (PRINT_POS_NAME ? params.out_position_names[g].second : std::to_string(params.out_position_names[g].first)) + ",";
        label_max = std::max(label_max, labels[g].size());
    }
    // A batch of consecutive samples at a time: (1) gather the batch's scores into a sample-major tile -- the score
    // vectors are target-major, so this is the one pass that walks 3 G separate heap vectors, done with contiguous
    // reads; (2) every thread formats whole samples into its own buffer; (3) the buffers go out with parallel pwrites at
    // offsets known from their sizes. The reference issues S x G hash lookups and one flush per row (eval/idash.cpp:436-469).
    const size_t batch = std::max<size_t>(1, std::min<size_t>(S, (size_t) host_threads() * 2));
    const size_t row_max = 12 + label_max + 3 * 16 + 1;      // "<sample>" + label + three "%g," + newline
    std::vector<float> tile(batch * G * 3);
    std::vector<std::vector<char>> bufs(batch);
    std::vector<size_t> used(batch, 0);
    for (auto &b : bufs) b.resize(G * row_max);
    for (size_t s0 = 0; s0 < S; s0 += batch) {
        const size_t nb = std::min(batch, S - s0);
        parallel_ranges(G, [&](size_t b, size_t e) {
            for (size_t g = b; g < e; ++g)
                for (int v = 0; v < 3; ++v) {
                    const float *src = (*rows[g])[v].data() + s0;
                    for (size_t i = 0; i < nb; ++i) tile[(i * G + g) * 3 + v] = src[i];
                }
        });
        parallel_ranges(nb, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                char *out = bufs[i].data();
                char sid[24];
                const size_t sid_n = (size_t) (std::to_chars(sid, sid + sizeof(sid), s0 + i).ptr - sid);
                const float *t = tile.data() + i * G * 3;
                for (size_t g = 0; g < G; ++g) {
                    memcpy(out, sid, sid_n); out += sid_n;
                    memcpy(out, labels[g].data(), labels[g].size()); out += labels[g].size();
                    // operator<<(float) = printf("%g", double): std::to_chars(general, precision 6) is specified as exactly that
                    out = std::to_chars(out, out + 24, (double) t[3 * g + 0], std::chars_format::general, 6).ptr; *out++ = ',';
                    out = std::to_chars(out, out + 24, (double) t[3 * g + 1], std::chars_format::general, 6).ptr; *out++ = ',';
                    out = std::to_chars(out, out + 24, (double) t[3 * g + 2], std::chars_format::general, 6).ptr; *out++ = '\n';
                }
                used[i] = (size_t) (out - bufs[i].data());
            }
        }, 1);
        std::vector<off_t> offs(nb);
        for (size_t i = 0; i < nb; ++i) { offs[i] = file_off; file_off += (off_t) used[i]; }
        std::atomic<bool> bad(false);
        parallel_ranges(nb, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                size_t lo = 0;
                while (lo < used[i]) {
                    const ssize_t r = ::pwrite(fd, bufs[i].data() + lo, used[i] - lo, offs[i] + (off_t) lo);
                    if (r < 0 && errno == EINTR) continue;
                    if (r <= 0) { bad = true; return; }
                    lo += (size_t) r;
                }
            }
        }, 1);
        REQUIRE_DRAMATICALLY(!bad, "short write to result file");
    }
    REQUIRE_DRAMATICALLY(::close(fd) == 0, "error closing result file");
}

// ---- the two GPU stages ----------------------------------------------------------------------------------------
int idash_host_device() {
    if (const char *e = getenv("IDASH_B200_DEVICE")) return atoi(e);
    if (const char *e = getenv("LOCAL_RANK")) return atoi(e);
    return 0;
}
double idash_host_last_gpu_seconds() { return g_last_gpu_seconds; }

void cloud_compute_score(EncryptedPredictions &enc_preds, const EncryptedData &enc_data, const Model &model, const IdashParams &params) {
    REQUIRE_DRAMATICALLY(params.k == 1, "blah");                                        // eval/idash.cpp:768
    REQUIRE_DRAMATICALLY(enc_preds.score.empty(), "shit happens again");                // createAndGet, eval/idash.h:181
    const uint64_t n_rows = model.model.size();
    const bool timing = getenv("IDASH_HOST_TIMING") != nullptr;
    const double t_begin = now_s();

    // Model -> CSR, rows in the map's iteration order (the order the reference walks them, eval/idash.cpp:772-777);
    // row pointers serially, entries by a pool of threads
    std::vector<uint32_t> out_bidx(n_rows);
    std::vector<uint64_t> row_ptr(n_rows + 1, 0);
    std::vector<const std::unordered_map<FeatBigIndex, int32_t> *> rows(n_rows);
    {
        uint64_t r = 0;
        for (const auto &row : model.model) {
            out_bidx[r] = row.first;
            rows[r] = &row.second;
            row_ptr[r + 1] = row_ptr[r] + row.second.size();
            ++r;
        }
    }
    std::vector<uint32_t> col(row_ptr[n_rows]);
    std::vector<int32_t> coef(row_ptr[n_rows]);
    std::atomic<bool> missing(false);
    parallel_ranges(n_rows, [&](size_t b, size_t e) {
        for (size_t r = b; r < e; ++r) {
            uint64_t k = row_ptr[r];
            for (const auto &c : *rows[r]) {
                if (c.first != params.constant_bigIndex() && enc_data.enc_data.find(params.feature_indexOf(c.first)) == enc_data.enc_data.end())
                    missing = true;
                col[k] = c.first; coef[k] = c.second;
                ++k;
            }
        }
    });
    REQUIRE_DRAMATICALLY(!missing, "shit happens before");                              // getTLWE, eval/idash.h:164

    const double t_csr0 = now_s();
    idash_b200_ctx *ctx = gpu_ctx();
    const double t0 = now_s();
    idash_b200_model_desc desc;
    desc.num_samples = params.NUM_SAMPLES; desc.num_regions = params.NUM_REGIONS; desc.region_size = params.REGION_SIZE;
    desc.n_rows = n_rows; desc.out_bidx = out_bidx.data(); desc.row_ptr = row_ptr.data(); desc.col = col.data(); desc.coef = coef.data();
    idash_b200_model *dm = nullptr;
    if (idash_b200_model_upload(ctx, &desc, &dm) != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_model_upload: " << idash_b200_last_error());
    const double t_upload = now_s();
    // outputs: one slab; the map is filled in the reference's insertion order, then record slot k goes to the k-th
    // element in ITERATION order, so that the slab is the image of encrypted_prediction.bin (eval/idash.cpp:607)
    auto slab = take_output_slab(n_rows);
    const double t_slab = now_s();
    for (uint64_t r = 0; r < n_rows; ++r) enc_preds.score.emplace(out_bidx[r], nullptr);
    std::vector<uint32_t> slot_of_row(n_rows);
    {
        // the k-th element in iteration order gets slot k; its row is found through a bigIndex -> row table
        std::unordered_map<FeatBigIndex, uint32_t> row_of;
        row_of.reserve(n_rows);
        for (uint64_t r = 0; r < n_rows; ++r) row_of.emplace(out_bidx[r], (uint32_t) r);
        uint64_t k = 0;
        for (auto &it : enc_preds.score) {
            it.second = &slab->samples[k];
            slot_of_row[row_of.at(it.first)] = (uint32_t) k;
            ++k;
        }
    }
    enc_preds.slab = slab;

    idash_b200_cts in, out;
    std::unique_ptr<PackedCts> staged;
    if (all_views_of(enc_data.enc_data, enc_data.slab)) {
        enc_data.slab->push_variances();
        in = {IDASH_B200_LAYOUT_RECORDS, enc_data.slab->records(), enc_data.slab->count, nullptr, nullptr};
    } else {
        staged = pack_map(enc_data.enc_data);
        in = {IDASH_B200_LAYOUT_PACKED, staged->words(), staged->count, staged->index(), staged->variance()};
    }
    out = {IDASH_B200_LAYOUT_RECORDS, slab->records(), n_rows, nullptr, nullptr};
    const double t_eval0 = now_s();
    const int rc = idash_b200_cloud_eval_host(ctx, dm, &in, &out, slot_of_row.data());
    if (rc != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_cloud_eval_host: " << idash_b200_last_error());
    const double t_eval1 = now_s();
    idash_b200_model_free(dm);
    g_last_gpu_seconds = (t_upload - t0) + (t_eval1 - t_eval0);
    slab->pull_variances();
    if (timing)
        fprintf(stderr, "[idash_host] cloud_compute_score: csr %.3f s, context wait %.3f s, model_upload %.3f s, output slab wait %.3f s, containers %.3f s, cloud_eval_host %.3f s, total %.3f s\n",
                t_csr0 - t_begin, t0 - t_csr0, t_upload - t0, t_slab - t_upload, t_eval0 - t_slab, t_eval1 - t_eval0, now_s() - t_begin);
}

void decrypt_predictions(DecryptedPredictions &predictions, const EncryptedPredictions &enc_preds, const IdashKey &key) {
    const IdashParams &params = *key.idashParams;
    const uint32_t S = params.NUM_SAMPLES;
    REQUIRE_DRAMATICALLY(key.tlweKey && key.tlweKey->params->N == (int32_t) IdashParams::N && key.tlweKey->params->k == 1,
                         "unsupported TLWE key");
    idash_b200_cts in;
    std::unique_ptr<PackedCts> staged;
    std::unordered_map<FeatBigIndex, uint64_t> slot;      // output bigIndex -> ciphertext slot in `in`
    slot.reserve(enc_preds.score.size());
    if (all_views_of(enc_preds.score, enc_preds.slab)) {
        in = {IDASH_B200_LAYOUT_RECORDS, enc_preds.slab->records(), enc_preds.slab->count, nullptr, nullptr};
        for (const auto &it : enc_preds.score) slot.emplace(it.first, (uint64_t) enc_preds.slab->slot_of(it.second));
    } else {
        staged = pack_map(enc_preds.score);
        in = {IDASH_B200_LAYOUT_PACKED, staged->words(), staged->count, staged->index(), staged->variance()};
        for (uint64_t i = 0; i < staged->count; ++i) slot.emplace(staged->index()[i], i);
    }
    const HostBuf sb = host_buf_alloc(std::max<size_t>(16, (size_t) in.count * S * sizeof(float)));
    float *scores = static_cast<float *>(sb.p);
    const double t0 = now_s();
    const int rc = idash_b200_decrypt_host(gpu_ctx(), key.tlweKey->key[0].coefs, S, &in, scores, nullptr);
    if (rc != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_decrypt_host: " << idash_b200_last_error());
    g_last_gpu_seconds = now_s() - t0;

    // fan out into score[pos][snp] (eval/idash.cpp:689-696, 714-719); a missing output dies like .at()
    struct Dst { std::array<std::vector<float>, 3> *v; std::array<uint64_t, 3> s; };
    std::vector<Dst> dst;
    dst.reserve(params.out_features_index.size());
    for (const auto &it : params.out_features_index) {
        auto &v = predictions.score[it.first];
        Dst d{&v, {0, 0, 0}};
        for (int snp = 0; snp < 3; ++snp) {
            const auto s = slot.find(it.second[snp]);
            REQUIRE_DRAMATICALLY(s != slot.end(), "encrypted predictions lack output feature " << it.second[snp]);
            d.s[snp] = s->second;
        }
        dst.push_back(d);
    }
    parallel_ranges(dst.size(), [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i)
            for (int snp = 0; snp < 3; ++snp) (*dst[i].v)[snp].assign(scores + dst[i].s[snp] * S, scores + (dst[i].s[snp] + 1) * S);
    });
    host_buf_free(sb);
}

void StageClock::print_benchmark(const char *stage_label, double stage_seconds) const {
    const double total = profiler.walltime();
    std::cout << "----------------- BENCHMARK ----------------- " << std::endl;
    std::cout << stage_label << stage_seconds << std::endl;
    std::cout << "serialization wall time (seconds): " << total - stage_seconds << std::endl;
    std::cout << "total wall time (seconds)........: " << total << std::endl;
    std::cout << "RAM usage (MB)...................: " << profiler.maxrss() / 1e6 << std::endl;
    std::cout << "gpu call wall time (seconds).....: " << idash_host_last_gpu_seconds() << std::endl;
}

// ---- Profiler (eval/idash.cpp:935-965) -------------------------------------------------------------------------
Profiler::Profiler() : tw0(universalWallTime()), tc0(universalClockTime()) {}
double Profiler::universalWallTime() {
    struct timeval tv;
    return gettimeofday(&tv, nullptr) ? 0.0 : (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec;
}
double Profiler::universalClockTime() { return (double) clock() / CLOCKS_PER_SEC; }
double Profiler::walltime() const { return universalWallTime() - tw0; }
double Profiler::clocktime() const { return universalClockTime() - tc0; }
long int Profiler::maxrss() const {
    struct rusage u;
    return getrusage(RUSAGE_SELF, &u) ? -1 : u.ru_maxrss * 1000;
}
