// Probe for the tcgen05 (kind::i8) operand conventions the K2 cloud kernel relies on. Runs a few
// single-tile MMAs on sm_100a and compares with a CPU product, so that the shared-memory descriptor
// fields (LBO / SBO for MN-major A and K-major B without swizzle), the mixed u8 x s8 instruction
// descriptor and the tcgen05.ld lane/column mapping are pinned by measurement, not by reading.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct MmaOp { uint32_t a_off, b_off, idesc, accumulate, d_col; };
struct ProbeArgs {
    const uint8_t *a_img; uint32_t a_bytes;
    const uint8_t *b_img; uint32_t b_bytes;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
    uint32_t n_ops;
    MmaOp ops[16];
    int32_t *out;        // [128][ncols]
    uint32_t ncols;      // TMEM columns to read back (multiple of 16)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t) ((addr >> 4) & 0x3FFF);
    d |= (uint64_t) ((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;   // version = 1 (Blackwell)
    return d;                  // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) umma_probe(const ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem, *sb = smem + 65536;
    for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += 128 * 16) *(uint4 *) (sa + i) = *(const uint4 *) (p.a_img + i);
    for (uint32_t i = threadIdx.x * 16; i < p.b_bytes; i += 128 * 16) *(uint4 *) (sb + i) = *(const uint4 *) (p.b_img + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < p.n_ops; ++i) {
            const uint64_t da = make_desc(smem_u32(sa) + p.ops[i].a_off, p.a_lbo, p.a_sbo);
            const uint64_t db = make_desc(smem_u32(sb) + p.ops[i].b_off, p.b_lbo, p.b_sbo);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tmem + p.ops[i].d_col), "l"(da), "l"(db), "r"(p.ops[i].idesc), "r"(p.ops[i].accumulate) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_row = warp * 32 + (threadIdx.x & 31);
    for (uint32_t c = 0; c < p.ncols; c += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((warp * 32u) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) p.out[lane_row * p.ncols + c + i] = (int32_t) v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

static uint32_t make_idesc(int M, int N, int a_signed, int b_signed, int a_mn_major, int b_mn_major) {
    return (2u << 4) | ((uint32_t) a_signed << 7) | ((uint32_t) b_signed << 10) | ((uint32_t) a_mn_major << 15) |
           ((uint32_t) b_mn_major << 16) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

// A[m][k] (MN-major, no swizzle): byte (m%16) + (k%8)*16 + (m/16)*mn_stride + (k/8)*k_stride
static void fill_a(std::vector<uint8_t> &img, const std::vector<uint8_t> &A, int M, int K, uint32_t mn_stride, uint32_t k_stride) {
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) img[(m % 16) + (k % 8) * 16 + (m / 16) * mn_stride + (k / 8) * k_stride] = A[m * K + k];
}
// B[n][k] (K-major, no swizzle): byte (k%16) + (n%8)*16 + (n/8)*mn_stride + (k/16)*k_stride
static void fill_b(std::vector<uint8_t> &img, const std::vector<uint8_t> &B, int N, int K, uint32_t mn_stride, uint32_t k_stride) {
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) img[(k % 16) + (n % 8) * 16 + (n / 8) * mn_stride + (k / 16) * k_stride] = B[n * K + k];
}

int main() {
    const int M = 128, N = 64;
    CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    uint8_t *d_a, *d_b; int32_t *d_out;
    CK(cudaMalloc(&d_a, 65536)); CK(cudaMalloc(&d_b, 32768)); CK(cudaMalloc(&d_out, 128 * 256 * 4));
    srand(7);
    int n_fail = 0;
    // variant: {K, a_mn_stride, a_k_stride, a_swap (put MN stride into LBO), b_signed, name}
    struct Var { int K; uint32_t a_mn, a_k; int a_swap; int b_signed; int two_acc; const char *name; };
    const Var vars[] = {
        {32, 128, 1024, 0, 0, 0, "K=32 A: SBO=MN stride 128, LBO=K stride 1024 (CUTLASS convention); B u8"},
        {32, 128, 1024, 1, 0, 0, "K=32 A: LBO=MN stride, SBO=K stride (swapped); B u8"},
        {32, 144, 1152, 0, 0, 0, "K=32 A padded: SBO=144, LBO=1152; B u8"},
        {32, 144, 1152, 0, 1, 0, "K=32 A padded u8 x B s8 (mixed signedness)"},
        {64, 144, 1152, 0, 1, 0, "K=64 as two MMAs, A u8 x B s8"},
        {64, 144, 1152, 0, 1, 1, "K=64, 4 accumulators at columns 0/64/128/192, u8 x u8 then u8 x s8 accumulated"},
    };
    for (const Var &v : vars) {
        const int K = v.K;
        std::vector<uint8_t> A(M * K), B(N * K), B2(N * K);
        for (auto &x : A) x = (uint8_t) rand();
        for (auto &x : B) x = (uint8_t) rand();
        for (auto &x : B2) x = (uint8_t) rand();
        const uint32_t b_mn = 128, b_k = N * 16;
        std::vector<uint8_t> a_img(65536, 0), b_img(32768, 0);
        fill_a(a_img, A, M, K, v.a_mn, v.a_k);
        fill_b(b_img, B, N, K, b_mn, b_k);
        const uint32_t b2_off = (K / 16) * b_k;
        { std::vector<uint8_t> tmp(32768, 0); fill_b(tmp, B2, N, K, b_mn, b_k); memcpy(b_img.data() + b2_off, tmp.data(), (K / 16) * b_k); }
        CK(cudaMemcpy(d_a, a_img.data(), 65536, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_b, b_img.data(), 32768, cudaMemcpyHostToDevice));
        ProbeArgs p; memset(&p, 0, sizeof(p));
        p.a_img = d_a; p.a_bytes = 65536; p.b_img = d_b; p.b_bytes = 32768;
        p.a_lbo = v.a_swap ? v.a_mn : v.a_k; p.a_sbo = v.a_swap ? v.a_k : v.a_mn;
        p.b_lbo = b_k; p.b_sbo = b_mn;
        p.out = d_out;
        std::vector<int64_t> ref((size_t) M * 256, 0);
        if (!v.two_acc) {
            p.ncols = 64;
            for (int ks = 0; ks < K / 32; ++ks) {
                MmaOp &o = p.ops[p.n_ops++];
                o.a_off = ks * 4 * v.a_k; o.b_off = ks * 2 * b_k; o.idesc = make_idesc(M, N, 0, v.b_signed, 1, 0); o.accumulate = ks > 0; o.d_col = 0;
            }
            for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                int64_t s = 0;
                for (int k = 0; k < K; ++k) s += (int64_t) A[m * K + k] * (v.b_signed ? (int64_t) (int8_t) B[n * K + k] : (int64_t) B[n * K + k]);
                ref[(size_t) m * 64 + n] = s;
            }
        } else {
            p.ncols = 256;
            for (int acc = 0; acc < 4; ++acc)
                for (int part = 0; part < 2; ++part)          // part 0: A x B (u8 x u8), part 1: A x B2 (u8 x s8)
                    for (int ks = 0; ks < K / 32; ++ks) {
                        MmaOp &o = p.ops[p.n_ops++];
                        o.a_off = ks * 4 * v.a_k; o.b_off = (part ? b2_off : 0) + ks * 2 * b_k;
                        o.idesc = make_idesc(M, N, 0, part, 1, 0); o.accumulate = (part | ks) > 0; o.d_col = acc * 64;
                    }
            for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                int64_t s = 0;
                for (int k = 0; k < K; ++k) s += (int64_t) A[m * K + k] * ((int64_t) B[n * K + k] + (int64_t) (int8_t) B2[n * K + k]);
                for (int acc = 0; acc < 4; ++acc) ref[(size_t) m * 256 + acc * 64 + n] = s;
            }
        }
        CK(cudaMemset(d_out, 0xCD, 128 * 256 * 4));
        umma_probe<<<1, 128, 98304>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL (%s): %s\n", cudaGetErrorString(e), v.name); return 1; }
        std::vector<int32_t> out((size_t) M * p.ncols);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < out.size(); ++i) bad += (int64_t) out[i] != ref[i];
        printf("%s: %s (%zu / %zu mismatches) sample out[1][2]=%d ref=%lld\n", bad ? "MISMATCH" : "MATCH", v.name, bad, out.size(),
               out[1 * p.ncols + 2], (long long) ref[1 * p.ncols + 2]);
        n_fail += bad != 0;
    }
    printf("probe done: %d variant(s) mismatched\n", n_fail);
    return 0;
}
