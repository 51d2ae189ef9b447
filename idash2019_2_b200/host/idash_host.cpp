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

#include <sys/stat.h>

#include "idash_b200.h"
#include "idash_b200_layout.h"

using std::string;

// The model as read_model emits it for the device (north-star: "parse_vw emits a device-resident block-banded layout"): rows sorted
// by output bigIndex (so that a contiguous target range is a contiguous row range: multi-GPU sharding and pipelined copies), one
// uploaded copy per GPU, and the order in which the reference would walk Model::model -- the record order of
// encrypted_prediction.bin (eval/idash.cpp:772-777, 607).
struct CompiledModel {
    uint32_t S = 0, NR = 0, RS = 0;
    uint64_t n_rows = 0;
    std::vector<uint32_t> sorted_bidx;      // device row -> output bigIndex, ascending
    std::vector<uint32_t> map_order;        // output bigIndices in the iteration order of the reference's Model::model
    bool from_cache = false;
    std::shared_future<std::vector<idash_b200_model *>> device_models;   // one per entry of idash_host_devices()
    ~CompiledModel() {
        if (device_models.valid())
            for (idash_b200_model *m : device_models.get()) idash_b200_model_free(m);
    }
};

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

// one context per GPU of idash_host_devices() (normally one: one process per GPU; IDASH_GPUS lists several)
// (thread-safe: the cloud binary creates them on a helper thread while the model files are parsed; `die` = false is that
// helper's probe -- a box without a GPU must still run the file-format half of this layer)
const std::vector<idash_b200_ctx *> &gpu_ctxs(bool die = true) {
    static std::vector<idash_b200_ctx *> ctxs;
    static string error;
    static std::once_flag once;
    std::call_once(once, []() {
        std::vector<int> devs = idash_host_devices();
        // CUDA initialisation enumerates and initialises every visible GPU (measured on an 8-GPU box: 6.7 s before the first context
        // exists, against 1.8 s with one GPU visible). Unless the caller has already restricted it, the process only exposes the GPUs
        // it is going to use; the runtime then numbers them 0 .. n-1. This must happen before the first CUDA call of the process.
        if (!getenv("CUDA_VISIBLE_DEVICES") && !getenv("IDASH_HOST_ALL_GPUS_VISIBLE")) {
            string list;
            for (size_t i = 0; i < devs.size(); ++i) { if (i) list += ','; list += std::to_string(devs[i]); }
            setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1);
            for (size_t i = 0; i < devs.size(); ++i) devs[i] = (int) i;
        }
        std::vector<idash_b200_ctx *> made(devs.size(), nullptr);
        std::vector<string> errs(devs.size());
        std::vector<std::thread> th;
        for (size_t i = 0; i < devs.size(); ++i)
            th.emplace_back([&, i]() { if (idash_b200_init(&made[i], devs[i]) != IDASH_B200_OK) { made[i] = nullptr; errs[i] = idash_b200_last_error(); } });
        for (auto &x : th) x.join();
        for (size_t i = 0; i < devs.size(); ++i)
            if (!made[i]) { error = errs[i]; for (idash_b200_ctx *c : made) if (c) idash_b200_destroy(c); return; }
        ctxs = made;
    });
    if (ctxs.empty() && die) DIE_DRAMATICALLY("idash_b200_init: " << error);
    return ctxs;
}
idash_b200_ctx *gpu_ctx(bool die = true) {
    const auto &c = gpu_ctxs(die);
    return c.empty() ? nullptr : c[0];
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

void warm_up_pin_input(const std::shared_ptr<CtSlab> &slab);

// whole ciphertext file -> slab; checks the size and every record's TLWE type uid (tfhe_io.cpp:308 aborts on a bad one).
// file_order (optional) receives, for the k-th record of the file, the slab slot it was put into.
// sort_by_index: the records are placed in the slab sorted by ciphertext index instead of file (hash) order, so that the inputs of a
// contiguous target range are one contiguous run of records -- what a GPU that owns a target range copies. Same single
// pass over the bytes: the file is mapped and every record is copied to its slot by all threads.
std::shared_ptr<CtSlab> read_ct_file(const string &filename, const char *what, bool sort_by_index, bool pin, std::vector<uint32_t> *file_order) {
    const int fd = ::open(filename.c_str(), O_RDONLY);
    REQUIRE_DRAMATICALLY(fd >= 0, "Cannot open encrypted " << what << " file for read");
    uint64_t count = 0;
    REQUIRE_DRAMATICALLY(::pread(fd, &count, 8, 0) == 8, "truncated encrypted " << what << " file");
    REQUIRE_DRAMATICALLY(count < ((uint64_t) 1 << 32), "encrypted " << what << " file: implausible record count");
    struct stat st;
    REQUIRE_DRAMATICALLY(::fstat(fd, &st) == 0, "cannot stat encrypted " << what << " file");
    // before anything is allocated: a corrupt count must end in this message, not in a 30 TB mapping
    REQUIRE_DRAMATICALLY(!S_ISREG(st.st_mode) || (uint64_t) st.st_size >= 8 + count * (uint64_t) REC, "truncated encrypted " << what << " file");
    auto slab = std::make_shared<CtSlab>(count);
    if (pin) warm_up_pin_input(slab);          // page-locking runs on a helper thread while the records are read
    const size_t bytes = (size_t) count * REC;
    void *map = MAP_FAILED;
    if (sort_by_index && count > 1 && !getenv("IDASH_HOST_NO_MMAP")) map = ::mmap(nullptr, 8 + bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map != MAP_FAILED) {
        const uint8_t *src = static_cast<const uint8_t *>(map) + 8;
        std::vector<uint32_t> idx(count);
        parallel_ranges(count, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) memcpy(&idx[i], src + i * REC, 4); });
        std::vector<uint32_t> by_index(count);
        for (uint64_t i = 0; i < count; ++i) by_index[i] = (uint32_t) i;
        std::stable_sort(by_index.begin(), by_index.end(), [&](uint32_t a, uint32_t b) { return idx[a] < idx[b]; });
        std::vector<uint32_t> slot_of(count);
        slab->sorted_index.resize(count);
        for (uint64_t s = 0; s < count; ++s) { slot_of[by_index[s]] = (uint32_t) s; slab->sorted_index[s] = idx[by_index[s]]; }
        parallel_ranges(count, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) memcpy(slab->record(slot_of[i]), src + i * REC, REC); });
        ::munmap(map, 8 + bytes);
        if (file_order) *file_order = std::move(slot_of);
    } else {
        parallel_file_io(fd, slab->records(), bytes, 8, false, what);
        if (file_order) { file_order->resize(count); for (uint64_t i = 0; i < count; ++i) (*file_order)[i] = (uint32_t) i; }
    }
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

template <class Map>
bool all_views_of(const Map &m, const std::shared_ptr<CtSlab> &slab);

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
    if (all_views_of(m, slab)) {
        // Samples are views into one slab but the map walks them in another order (cloud_compute_score keeps its output slab in
        // target order and the reference's record order is hash order): every record is copied from its slab slot to its file
        // slot, by all threads, straight into the mapped file -- the same single pass over the bytes as the bulk copy above.
        const uint64_t count = m.size();
        std::vector<std::pair<uint32_t, const TLweSample *>> order;
        order.reserve(count);
        for (const auto &it : m) order.push_back({it.first, it.second});
        const size_t bytes = 8 + (size_t) count * REC;
        const int fd = ::open(filename.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        REQUIRE_DRAMATICALLY(fd >= 0, "Cannot open encrypted " << what << " file for write");
        REQUIRE_DRAMATICALLY(::ftruncate(fd, (off_t) bytes) == 0, "cannot size encrypted " << what << " file");
        void *map = getenv("IDASH_HOST_NO_MMAP") ? MAP_FAILED : ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        const int32_t uid = IDASH_B200_TLWE_SAMPLE_UID;
        auto fill = [&](uint8_t *dst, uint64_t k) {
            memcpy(dst, slab->record((uint64_t) slab->slot_of(order[k].second)), REC);
            memcpy(dst, &order[k].first, 4);
            memcpy(dst + 4, &uid, 4);
            memcpy(dst + REC_VAR_OFF, &order[k].second->current_variance, 8);
        };
        std::atomic<bool> bad(false);
        if (map != MAP_FAILED) {
            memcpy(map, &count, 8);
            parallel_ranges(count, [&](size_t b, size_t e) { for (size_t k = b; k < e; ++k) fill(static_cast<uint8_t *>(map) + 8 + k * REC, k); });
            REQUIRE_DRAMATICALLY(::munmap(map, bytes) == 0, "error unmapping encrypted " << what << " file");
        } else {
            REQUIRE_DRAMATICALLY(::pwrite(fd, &count, 8, 0) == 8, "short write to encrypted " << what << " file");
            parallel_ranges(count, [&](size_t b, size_t e) {
                const size_t batch = 256;
                std::vector<uint8_t> buf(batch * REC);
                for (size_t k0 = b; k0 < e; k0 += batch) {
                    const size_t n = std::min(batch, e - k0);
                    for (size_t i = 0; i < n; ++i) fill(buf.data() + i * REC, k0 + i);
                    size_t lo = 0;
                    while (lo < n * REC) {
                        const ssize_t r = ::pwrite(fd, buf.data() + lo, n * REC - lo, (off_t) (8 + k0 * REC + lo));
                        if (r < 0 && errno == EINTR) continue;
                        if (r <= 0) { bad = true; return; }
                        lo += (size_t) r;
                    }
                }
            });
        }
        REQUIRE_DRAMATICALLY(!bad, "short write to encrypted " << what << " file");
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

// Helper-thread warm-up for the cloud stage (started by read_params / read_model / read_encrypted_data, collected by
// idash_host_wait_ready and cloud_compute_score): GPU contexts, the page-locked slab the output ciphertexts are copied into,
// page-locking of the input slab, upload of the compiled model.
std::mutex g_warm_mutex;
std::shared_future<void> g_warm_ctx;
std::shared_future<std::shared_ptr<CtSlab>> g_warm_slab;
std::vector<std::shared_future<void>> g_warm_misc;
string g_model_cache_path;

// page-lock a slab once the context exists (the C ABI copies to / from page-locked memory at the link's full rate and
// asynchronously; measured on the 2 GB output slab: 0.036 s against 0.105 s pageable). Best effort.
void pin_slab(const std::shared_ptr<CtSlab> &slab) {
    if (!slab || slab->count == 0 || slab->mem_kind == HOST_MEM_PINNED || getenv("IDASH_HOST_NO_PIN")) return;
    if (!gpu_ctx(false)) return;
    if (idash_b200_host_register(slab->mem, slab->mem_bytes) == IDASH_B200_OK) slab->registered = true;
}

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
    std::shared_future<void> ctx_ready = g_warm_ctx;
    g_warm_slab = std::async(std::launch::async, [n_rows, ctx_ready]() {
        const double t0 = now_s();
        auto slab = std::make_shared<CtSlab>(n_rows);
        const double t1 = now_s();
        if (ctx_ready.valid()) ctx_ready.wait();
        pin_slab(slab);
        if (getenv("IDASH_HOST_TIMING"))
            fprintf(stderr, "[idash_host] warm-up: output slab (%.2f GB) allocated in %.3f s, page-locked %s after %.3f s\n", slab->image_bytes() * 1e-9,
                    t1 - t0, slab->registered ? "yes" : "no", now_s() - t0);
        return slab;
    }).share();
}
void warm_up_pin_input(const std::shared_ptr<CtSlab> &slab) {
    if (getenv("IDASH_HOST_NO_WARMUP")) return;
    std::lock_guard<std::mutex> lock(g_warm_mutex);
    if (!g_warm_ctx.valid()) return;             // no GPU stage in sight (file-format use of this layer)
    std::shared_future<void> ctx_ready = g_warm_ctx;
    g_warm_misc.push_back(std::async(std::launch::async, [slab, ctx_ready]() { ctx_ready.wait(); pin_slab(slab); }).share());
}
std::shared_ptr<CtSlab> take_output_slab(uint64_t n_rows) {
    std::shared_ptr<CtSlab> slab;
    {
        std::lock_guard<std::mutex> lock(g_warm_mutex);
        if (g_warm_slab.valid()) { slab = g_warm_slab.get(); g_warm_slab = std::shared_future<std::shared_ptr<CtSlab>>(); }
    }
    if (!slab || slab->count != n_rows) { slab = std::make_shared<CtSlab>(n_rows); pin_slab(slab); }
    return slab;
}

uint64_t fnv1a(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Uploads a compiled layout to every GPU of this process (helper thread; the layout is shared between the copies).
std::shared_future<std::vector<idash_b200_model *>> upload_everywhere(idash_b200_layout *layout) {
    std::shared_future<void> ctx_ready;
    {
        std::lock_guard<std::mutex> lock(g_warm_mutex);
        ctx_ready = g_warm_ctx;
    }
    return std::async(std::launch::async, [layout, ctx_ready]() {
        if (ctx_ready.valid()) ctx_ready.wait();
        const double t0 = now_s();
        const auto &ctxs = gpu_ctxs(false);
        std::vector<idash_b200_model *> ms(ctxs.size(), nullptr);
        if (ctxs.empty()) { idash_b200_layout_free(layout); return ms; }       // no GPU: file-format use of this layer; the cloud stage dies later
        if (idash_b200_model_upload_layout(ctxs[0], layout, &ms[0]) != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_model_upload_layout: " << idash_b200_last_error());
        for (size_t g = 1; g < ctxs.size(); ++g)
            if (idash_b200_model_clone(ctxs[g], ms[0], &ms[g]) != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_model_clone: " << idash_b200_last_error());
        if (getenv("IDASH_HOST_TIMING")) fprintf(stderr, "[idash_host] model uploaded to %zu GPU(s) in %.3f s\n", ctxs.size(), now_s() - t0);
        return ms;
    }).share();
}
void track_upload(const std::shared_future<std::vector<idash_b200_model *>> &f) {
    std::lock_guard<std::mutex> lock(g_warm_mutex);
    g_warm_misc.push_back(std::async(std::launch::deferred, [f]() { f.wait(); }).share());
}

// CSR (rows sorted by output bigIndex) -> compiled layout (+ cache file) -> CompiledModel with the upload in flight
std::shared_ptr<CompiledModel> compile_model(const IdashParams &params, std::vector<uint32_t> &&sorted_bidx, const std::vector<uint64_t> &row_ptr,
                                             const std::vector<uint32_t> &col, const std::vector<int32_t> &coef, std::vector<uint32_t> &&map_order,
                                             const string &cache_path, uint64_t cache_key) {
    auto cm = std::make_shared<CompiledModel>();
    cm->S = params.NUM_SAMPLES; cm->NR = params.NUM_REGIONS; cm->RS = params.REGION_SIZE;
    cm->n_rows = sorted_bidx.size();
    idash_b200_model_desc desc;
    desc.num_samples = cm->S; desc.num_regions = cm->NR; desc.region_size = cm->RS;
    desc.n_rows = cm->n_rows; desc.out_bidx = sorted_bidx.data(); desc.row_ptr = row_ptr.data(); desc.col = col.data(); desc.coef = coef.data();
    idash_b200_layout *layout = nullptr;
    if (idash_b200_layout_compile_ex(&desc, IDASH_B200_COMPILE_DEFAULT, &layout) != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_layout_compile: " << idash_b200_last_error());
    if (!cache_path.empty() && idash_b200_layout_save(layout, cache_path.c_str(), cache_key) != IDASH_B200_OK && getenv("IDASH_HOST_TIMING"))
        fprintf(stderr, "[idash_host] model cache not written: %s\n", idash_b200_last_error());
    cm->sorted_bidx = std::move(sorted_bidx);
    cm->map_order = std::move(map_order);
    cm->device_models = upload_everywhere(layout);
    track_upload(cm->device_models);
    return cm;
}

// Model (map of maps) -> CSR with rows sorted by output bigIndex -> compiled model. Row pointers serially, entries by all threads.
std::shared_ptr<CompiledModel> compile_from_map(const Model &model, const IdashParams &params, const string &cache_path, uint64_t cache_key) {
    const uint64_t n_rows = model.model.size();
    std::vector<uint32_t> map_order;
    map_order.reserve(n_rows);
    std::vector<std::pair<uint32_t, const std::unordered_map<FeatBigIndex, int32_t> *>> rows;
    rows.reserve(n_rows);
    for (const auto &row : model.model) { map_order.push_back(row.first); rows.push_back({row.first, &row.second}); }
    std::sort(rows.begin(), rows.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    std::vector<uint32_t> sorted_bidx(n_rows);
    std::vector<uint64_t> row_ptr(n_rows + 1, 0);
    for (uint64_t r = 0; r < n_rows; ++r) { sorted_bidx[r] = rows[r].first; row_ptr[r + 1] = row_ptr[r] + rows[r].second->size(); }
    std::vector<uint32_t> col(row_ptr[n_rows]);
    std::vector<int32_t> coef(row_ptr[n_rows]);
    parallel_ranges(n_rows, [&](size_t b, size_t e) {
        for (size_t r = b; r < e; ++r) {
            uint64_t k = row_ptr[r];
            for (const auto &c : *rows[r].second) { col[k] = c.first; coef[k] = c.second; ++k; }
        }
    });
    return compile_model(params, std::move(sorted_bidx), row_ptr, col, coef, std::move(map_order), cache_path, cache_key);
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
CtSlab::~CtSlab() {
    if (registered) idash_b200_host_unregister(mem);
    host_buf_free(HostBuf{mem, mem_kind, mem_bytes});
}
uint32_t CtSlab::index_of(uint64_t i) const { uint32_t v; memcpy(&v, record(i), 4); return v; }
void CtSlab::pull_variances() {
    parallel_ranges(count, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) memcpy(&samples[i].current_variance, record(i) + REC_VAR_OFF, 8); }, 4096);
}
void CtSlab::push_variances() {
    parallel_ranges(count, [&](size_t b, size_t e) { for (size_t i = b; i < e; ++i) memcpy(record(i) + REC_VAR_OFF, &samples[i].current_variance, 8); }, 4096);
}

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

// eval/idash.cpp:510-533. The map is filled in FILE order (emplace: the first record of an index wins), as the reference does -- its
// iteration order is what write_encrypted_data reproduces; the slab itself holds the records sorted by index (read_ct_file).
void read_encrypted_data(EncryptedData &d, const IdashParams &, const string &filename) {
    std::vector<uint32_t> file_order;
    d.slab = read_ct_file(filename, "data", true, true, &file_order);
    for (uint64_t k = 0; k < d.slab->count; ++k) d.enc_data.emplace(d.slab->index_of(file_order[k]), &d.slab->samples[file_order[k]]);
}

void read_encrypted_predictions(EncryptedPredictions &p, const IdashParams &, const string &filename) {
    p.slab = read_ct_file(filename, "predictions", false, false, nullptr);
    for (uint64_t i = 0; i < p.slab->count; ++i) p.score.emplace(p.slab->index_of(i), &p.slab->samples[i]);
}

void write_encrypted_data(const EncryptedData &d, const IdashParams &, const string &filename) {
    write_ct_file(d.enc_data, d.slab, filename, "data");
}

void write_encrypted_predictions(const EncryptedPredictions &p, const IdashParams &, const string &filename) {
    write_ct_file(p.score, p.slab, filename, "predictions");
}

// The order in which the reference walks Model::model (eval/idash.cpp:772-777): keys inserted in the iteration order of
// out_features_index, variants 0..2 (eval/idash.cpp:75-88). The iteration order of a libstdc++ unordered_map depends on the key type,
// the hasher, the insertion sequence and the rehash policy only -- not on the mapped type -- so a map of the same keys to a byte
// reproduces it without the coefficient maps (what the model-cache path needs).
static std::vector<uint32_t> reference_row_order(const IdashParams &params) {
    std::unordered_map<FeatBigIndex, char> probe;
    for (const auto &e : params.out_features_index)
        for (int snp = 0; snp < 3; ++snp) probe[e.second[(size_t) snp]] = 0;
    std::vector<uint32_t> order;
    order.reserve(probe.size());
    for (const auto &it : probe) order.push_back(it.first);
    return order;
}

// Fingerprint of what a compiled model depends on: the parameters (sizes and both feature index tables) and the model directory
// (name, size and mtime of every <pos>_<variant>.hr file, combined order-independently; stat'ed by all threads: 0.02-0.05 s for
// 242 646 files against 0.25 s for parsing them). false = some file is missing (read_model then dies like the reference).
static bool model_fingerprint(const IdashParams &params, const string &path, uint64_t *key) {
    uint64_t h = 1469598103934665603ull;
    const uint32_t head[7] = {params.NUM_SAMPLES, params.NUM_INPUT_POSITIONS, params.NUM_OUTPUT_POSITIONS, params.NUM_INPUT_FEATURES,
                              params.NUM_OUTPUT_FEATURES, params.NUM_REGIONS, params.REGION_SIZE};
    h = fnv1a(h, head, sizeof(head));
    uint64_t tags = 0;                                         // order-independent: the map's iteration order is not part of the model
    for (const auto &e : params.in_features_index) tags += fnv1a(fnv1a(1469598103934665603ull, &e.first, 8), e.second.data(), 12);
    h = fnv1a(h, &tags, 8);
    std::vector<std::pair<uint64_t, std::array<FeatBigIndex, 3>>> outs(params.out_features_index.begin(), params.out_features_index.end());
    std::atomic<uint64_t> files(0);
    std::atomic<bool> missing(false);
    parallel_ranges(outs.size(), [&](size_t b, size_t e) {
        uint64_t acc = 0;
        string fn;
        for (size_t i = b; i < e; ++i) {
            for (int snp = 0; snp < 3; ++snp) {
                fn = path; fn += '/'; fn += std::to_string(outs[i].first); fn += '_'; fn += (char) ('0' + snp); fn += ".hr";
                struct stat st;
                if (::stat(fn.c_str(), &st) != 0) { missing = true; return; }
                const uint64_t rec[5] = {outs[i].first, (uint64_t) outs[i].second[(size_t) snp], (uint64_t) st.st_size, (uint64_t) st.st_mtim.tv_sec, (uint64_t) st.st_mtim.tv_nsec};
                acc += fnv1a(1469598103934665603ull, rec, sizeof(rec));
            }
        }
        files += acc;
    });
    if (missing) return false;
    const uint64_t f = files.load();
    h = fnv1a(h, &f, 8);
    *key = h;
    return true;
}

// eval/idash.cpp:66-90. The .hr files are parsed by a pool of threads (3 x NUM_OUTPUT_POSITIONS small files);
// the model map is then filled in the reference's order (iteration order of out_features_index, variant 0..2),
// which fixes the record order of encrypted_prediction.bin downstream. The model is compiled into the device layout right here
// (north-star: "parse_vw emits a device-resident block-banded layout") and its upload starts on a helper thread; with a model cache
// (idash_host_set_model_cache) a later run takes the compiled layout from one file and skips the text files altogether.
void read_model(Model &model, const IdashParams &params, const string &path) {
    // Only the cloud stage loads a model: bring the GPU context up and allocate + fault in the output slab on helper
    // threads while the files are parsed (CUDA initialisation and 2 GB of fresh host memory are the two largest fixed
    // costs of cloud_compute_score at iDASH scale).
    warm_up_cloud(3 * (uint64_t) params.out_features_index.size());
    const bool timing = getenv("IDASH_HOST_TIMING") != nullptr;
    const double t_begin = now_s();

    string cache_path;
    {
        std::lock_guard<std::mutex> lock(g_warm_mutex);
        cache_path = g_model_cache_path;
    }
    uint64_t cache_key = 0;
    if (!cache_path.empty()) {
        if (!model_fingerprint(params, path, &cache_key)) cache_path.clear();      // a missing file: parse, and die on it like the reference
        idash_b200_layout *layout = nullptr;
        if (!cache_path.empty() && idash_b200_layout_load(cache_path.c_str(), cache_key, &layout) == IDASH_B200_OK) {
            idash_b200_model_info info;
            idash_b200_layout_get_info(layout, &info);
            auto cm = std::make_shared<CompiledModel>();
            cm->S = params.NUM_SAMPLES; cm->NR = params.NUM_REGIONS; cm->RS = params.REGION_SIZE;
            cm->n_rows = info.n_rows;
            cm->sorted_bidx.assign(idash_b200_layout_out_bidx(layout), idash_b200_layout_out_bidx(layout) + info.n_rows);
            cm->map_order = reference_row_order(params);
            cm->from_cache = true;
            if (cm->map_order.size() == cm->n_rows) {
                cm->device_models = upload_everywhere(layout);
                track_upload(cm->device_models);
                model.compiled = cm;
                if (timing) fprintf(stderr, "[idash_host] read_model: compiled model taken from %s in %.3f s\n", cache_path.c_str(), now_s() - t_begin);
                return;
            }
            idash_b200_layout_free(layout);                    // not this model after all: parse
        }
    }

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
    const double t_filled = now_s();
    model.compiled = compile_from_map(model, params, cache_path, cache_key);
    if (timing)
        fprintf(stderr, "[idash_host] read_model: parse %.3f s, fill %.3f s, compile%s %.3f s\n", t_parsed - t_begin, t_filled - t_parsed,
                cache_path.empty() ? "" : " + cache write", now_s() - t_filled);
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
    const size_t row_max = 12 + label_max + 3 * 16 + 1;      // "<sample>" + label + three "%g," + newline
    // samples per batch: two per thread, but at most ~1 GB of row buffers (a 192-thread host would otherwise allocate 2.6 GB here)
    const size_t by_budget = std::max<size_t>(1, ((size_t) 1 << 30) / std::max<size_t>(1, G * row_max));
    const size_t batch = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(S, (size_t) host_threads() * 2), by_budget));
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
std::vector<int> idash_host_devices() {
    std::vector<int> devs;
    if (const char *e = getenv("IDASH_GPUS")) {
        for (const char *p = e; *p;) {
            char *end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            if (v >= 0 && std::find(devs.begin(), devs.end(), (int) v) == devs.end()) devs.push_back((int) v);
            p = end;
            while (*p == ',' || *p == ' ') ++p;
        }
    }
    if (devs.empty()) devs.push_back(idash_host_device());
    return devs;
}
double idash_host_last_gpu_seconds() { return g_last_gpu_seconds; }

void idash_host_set_model_cache(const string &path) {
    std::lock_guard<std::mutex> lock(g_warm_mutex);
    g_model_cache_path = path;
}

void idash_host_wait_ready() {
    std::shared_future<void> ctx;
    std::shared_future<std::shared_ptr<CtSlab>> slab;
    std::vector<std::shared_future<void>> misc;
    {
        std::lock_guard<std::mutex> lock(g_warm_mutex);
        ctx = g_warm_ctx; slab = g_warm_slab; misc.swap(g_warm_misc);
    }
    if (ctx.valid()) ctx.wait();
    if (slab.valid()) slab.wait();
    for (auto &f : misc) f.wait();
}

// eval/idash.cpp:763-848. No arithmetic here: the compiled model (read_model) is evaluated by libidash_b200 on one GPU, or -- with
// IDASH_GPUS -- on several, each taking a contiguous target range (rows sorted by output bigIndex, cut at tile boundaries) and only
// the part of the input slab its windows touch (the loop being cut is eval/idash.cpp:779-790). The output slab is kept in
// target order, which makes every GPU's share one contiguous device->host copy; the reference's record order is restored by
// write_encrypted_predictions while the records go to the file (one pass over the bytes either way).
void cloud_compute_score(EncryptedPredictions &enc_preds, const EncryptedData &enc_data, const Model &model, const IdashParams &params) {
    REQUIRE_DRAMATICALLY(params.k == 1, "blah");                                        // eval/idash.cpp:768
    REQUIRE_DRAMATICALLY(enc_preds.score.empty(), "shit happens again");                // createAndGet, eval/idash.h:181
    const bool timing = getenv("IDASH_HOST_TIMING") != nullptr;
    const double t_begin = now_s();

    std::shared_ptr<CompiledModel> cm = model.compiled;
    if (!cm || cm->S != params.NUM_SAMPLES || cm->NR != params.NUM_REGIONS || cm->RS != params.REGION_SIZE ||
        (!cm->from_cache && cm->n_rows != model.model.size())) {
        warm_up_context();
        cm = compile_from_map(model, params, string(), 0);                              // a hand-built / edited Model
    }
    const uint64_t n_rows = cm->n_rows;
    const double t_model = now_s();
    const auto &ctxs = gpu_ctxs();                                                       // dies without a GPU
    const std::vector<idash_b200_model *> &dms = cm->device_models.get();
    REQUIRE_DRAMATICALLY(dms.size() == ctxs.size() && !dms.empty(), "the compiled model was not uploaded");
    const double t_ready = now_s();
    auto slab = take_output_slab(n_rows);
    const double t_slab = now_s();

    idash_b200_cts in, out;
    std::unique_ptr<PackedCts> staged;
    const bool in_views = all_views_of(enc_data.enc_data, enc_data.slab);
    if (in_views) {
        enc_data.slab->push_variances();
        in = {IDASH_B200_LAYOUT_RECORDS, enc_data.slab->records(), enc_data.slab->count, nullptr, nullptr};
    } else {
        staged = pack_map(enc_data.enc_data);
        in = {IDASH_B200_LAYOUT_PACKED, staged->words(), staged->count, staged->index(), staged->variance()};
    }
    out = {IDASH_B200_LAYOUT_RECORDS, slab->records(), n_rows, nullptr, nullptr};

    // the GPU work runs on helper threads (one per GPU) while this thread builds the output container
    idash_b200_model_info info;
    if (idash_b200_model_get_info(dms[0], &info) != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_model_get_info: " << idash_b200_last_error());
    const size_t n_gpus = (ctxs.size() > 1 && info.ring_ok && info.n_overflow_rows == 0 && info.n_tiles >= 2 * ctxs.size()) ? ctxs.size() : 1;
    const double t_eval0 = now_s();
    std::vector<std::future<std::pair<int, string>>> work;
    for (size_t g = 0; g < n_gpus; ++g) {
        work.push_back(std::async(std::launch::async, [&, g]() -> std::pair<int, string> {
            int rc;
            if (n_gpus == 1) {
                rc = idash_b200_cloud_eval_host(ctxs[0], dms[0], &in, &out, nullptr);
            } else {
                const uint64_t T = info.n_tiles;
                const uint64_t rb = (T * g / n_gpus) * IDASH_B200_TILE_ROWS, re = std::min<uint64_t>(n_rows, (T * (g + 1) / n_gpus) * IDASH_B200_TILE_ROWS);
                idash_b200_cts part = in;
                const auto &sidx = enc_data.slab ? enc_data.slab->sorted_index : std::vector<uint32_t>();
                uint32_t ct_lo = 0, ct_hi = 0;
                if (in_views && !sidx.empty() && idash_b200_model_input_range(dms[g], rb, re, &ct_lo, &ct_hi) == IDASH_B200_OK) {
                    // the slab is sorted by ciphertext index: this GPU's share of the inputs is one contiguous run of records
                    const size_t lo = (size_t) (std::lower_bound(sidx.begin(), sidx.end(), ct_lo) - sidx.begin());
                    const size_t hi = (size_t) (std::lower_bound(sidx.begin(), sidx.end(), ct_hi) - sidx.begin());
                    part.data = enc_data.slab->record(lo);
                    part.count = hi - lo;
                }
                rc = idash_b200_cloud_eval_host_rows(ctxs[g], dms[g], &part, &out, rb, re);
            }
            return {rc, rc == IDASH_B200_OK ? string() : string(idash_b200_last_error())};
        }));
    }
    // outputs: one ciphertext per model row, inserted in the order the reference walks its model (eval/idash.cpp:772-777), so that
    // the map's iteration order -- the record order of encrypted_prediction.bin (eval/idash.cpp:607) -- is the reference's
    {
        const std::vector<uint32_t> &sb = cm->sorted_bidx;
        for (const uint32_t bidx : cm->map_order) {
            const size_t r = (size_t) (std::lower_bound(sb.begin(), sb.end(), bidx) - sb.begin());
            enc_preds.score.emplace(bidx, &slab->samples[r]);
        }
    }
    enc_preds.slab = slab;
    const double t_cont = now_s();
    int rc = IDASH_B200_OK;
    string err;
    for (auto &w : work) { const auto r = w.get(); if (r.first != IDASH_B200_OK && rc == IDASH_B200_OK) { rc = r.first; err = r.second; } }
    REQUIRE_DRAMATICALLY(rc != IDASH_B200_ERR_MISSING_INPUT, "shit happens before");     // getTLWE, eval/idash.h:164
    if (rc != IDASH_B200_OK) DIE_DRAMATICALLY("idash_b200_cloud_eval_host: " << err);
    const double t_eval1 = now_s();
    g_last_gpu_seconds = t_eval1 - t_eval0;
    slab->pull_variances();
    if (timing)
        fprintf(stderr, "[idash_host] cloud_compute_score: model %.3f s, wait for context + model upload %.3f s, output slab wait %.3f s, "
                        "evaluation on %zu GPU(s) %.3f s (containers built meanwhile in %.3f s), total %.3f s\n",
                t_model - t_begin, t_ready - t_model, t_slab - t_ready, n_gpus, t_eval1 - t_eval0, t_cont - t_eval0, now_s() - t_begin);
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
