// tcgen05.mma kind::i8 issue / execution rate for the operand layouts of the cloud kernels, and the latency of the
// commit -> mbarrier -> wait round trip. One CTA per SM; thread 0 issues `nm` MMAs (M = 128, N = n, K = 32), commits,
// waits, `rounds` times; clock64 around the loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t) 1 << 46) | ((uint64_t) swz << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (2u << 4) | (0u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct Args {
    uint32_t a_lbo, a_sbo, a_mn, a_plane, swz;   // A descriptor strides, MN-major flag, bytes between the 4 limb planes
    uint32_t b_lbo, b_sbo;
    uint32_t n;                             // N of the MMAs
    uint32_t nm, rounds;
    uint32_t same_acc;                      // 0: the cloud kernel's order (plane 0, 2, 1, 3 per K step, overlapping P0..P3 columns)
                                            // 1: all MMAs accumulate into the same columns
                                            // 2: grouped by plane (all K steps of plane 0, then 2, 1, 3): same columns back to back
                                            // 3: like 0 but four disjoint accumulators (columns 0, 128, 256, 384)
    uint32_t wait_each;                     // 1: wait for the commit every round; 0: only at the end (pure issue + execution rate)
    unsigned long long *cycles;             // per CTA
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        const uint32_t idesc = make_idesc(128, p.n, p.a_mn, 0);
        uint32_t phase = 0;
        const uint64_t da0 = make_desc(a0, p.a_lbo, p.a_sbo, p.swz), db0 = make_desc(b0, p.swz ? 16u : p.b_lbo, p.swz ? 256u : p.b_sbo, p.swz);
        const uint64_t pl = p.a_plane >> 4;
        uint32_t dcols[4] = {0u, 128u, 64u, 192u};
        if (p.same_acc == 1u || p.n > 128u) dcols[1] = dcols[2] = dcols[3] = 0u;
        if (p.same_acc == 3u) { dcols[1] = 128u; dcols[2] = 256u; dcols[3] = 384u; }
        const long long t0 = clock64();
        for (uint32_t r = 0; r < p.rounds; ++r) {
            // descriptors precomputed; 4 MMAs per K step, fully unrolled, K steps walk 2 blocks
            for (uint32_t ks = 0; ks < p.nm / 4u; ++ks) {
                const uint64_t da = da0 + (uint64_t) ((ks & 1u) * ((4u * p.a_plane) >> 4));
                const uint64_t db = db0 + (uint64_t) ((ks & 1u) * (4096u >> 4));
                const uint32_t acc = ks > 0;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, q;\n\t}\n"
                                 ::"r"(tmem + dcols[j]), "l"(da + j * pl), "l"(db), "r"(idesc), "r"(j >= 2 ? 1u : acc) : "memory");
                }
            }
            if (p.wait_each || r + 1 == p.rounds) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                phase ^= 1u;
            }
        }
        p.cycles[blockIdx.x] = (unsigned long long) (clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int smem = 160 * 1024;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long *d_cyc, h_cyc[256];
    CK(cudaMalloc(&d_cyc, 256 * 8));
    struct Lay { const char *name; uint32_t lbo, sbo, mn, plane, swz; };
    const Lay lays[] = {
        {"A MN-major padded (SBO 144, LBO 1152)", 1152, 144, 1, 4 * 1152, 0},
        {"A and B K-major SWIZZLE_32B (SBO 256)", 16, 256, 0, 4096, 6},
    };
    for (int grid : {prop.multiProcessorCount}) {
        for (const Lay &l : lays) {
            for (uint32_t n : {128u, 64u, 256u}) {
              for (uint32_t order : {0u, 1u, 3u}) {
                if (n == 256u && order != 1u) continue;
                for (uint32_t wait_each : {1u, 0u}) {
                    printf("grid %3d  %-40s N=%3u order %u %s:", grid, l.name, n, order, wait_each ? "commit+wait/round" : "no waits         ");
                    for (uint32_t nm : {0u, 4u, 8u, 16u, 24u}) {
                        if (!wait_each && nm == 0) { printf("      -"); continue; }
                        Args a;
                        a.a_lbo = l.lbo; a.a_sbo = l.sbo; a.a_mn = l.mn; a.a_plane = l.plane; a.swz = l.swz;
                        a.b_lbo = n == 256 ? 4096 : 2048; a.b_sbo = 128; a.n = n; a.nm = nm; a.rounds = 200; a.same_acc = order; a.wait_each = wait_each;
                        a.cycles = d_cyc;
                        rate_kernel<<<grid, 128, smem>>>(a);
                        CK(cudaDeviceSynchronize());
                        CK(cudaMemcpy(h_cyc, d_cyc, 8 * grid, cudaMemcpyDeviceToHost));
                        unsigned long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = h_cyc[i] > mx ? h_cyc[i] : mx;
                        printf(" nm=%2u %6.0f", nm, (double) mx / a.rounds);
                    }
                    printf("  cycles/round\n");
                }
              }
            }
        }
    }
    return 0;
}
