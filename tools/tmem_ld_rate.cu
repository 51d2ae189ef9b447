// tcgen05.ld throughput by load width, alone and against running tcgen05.mma (kind::i8, M = 128, N = 128): how much TMEM
// bandwidth the epilogue warps of the ring kernel have, and what the accumulator read-modify-writes of concurrent MMAs take away.
// One CTA per SM, 8 reader warps (two per TMEM lane quadrant, like the ring kernel's epilogue) + 1 MMA warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_rate tools/tmem_ld_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t) 1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
    return (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct Args {
    uint32_t width;      // columns per tcgen05.ld: 8, 16, 32, 64
    uint32_t rounds;     // each round: every reader warp reads 256 columns (one "tile" of 4 accumulators x 64 columns)
    uint32_t mma;        // 1: the MMA warp issues N = 128 MMAs back to back while the readers run
    uint32_t readers;    // 0: no loads (MMA rate alone)
    unsigned long long *cycles;   // [grid][2]: readers' cycles, MMAs issued by the MMA warp
};

template <int W> __device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t ld_cols<8>(uint32_t taddr) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    return s;
}
template <> __device__ __forceinline__ uint32_t ld_cols<16>(uint32_t taddr) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    return s;
}
template <> __device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    return s;
}

template <int W>
__global__ void __launch_bounds__(288, 1) rate_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile uint32_t stop_s;
    for (uint32_t i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        stop_s = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    unsigned long long n_mma = 0;
    if (warp < 8) {
        // reader warp: lane quadrant warp % 4, columns 32 (warp / 4) .. of each of 4 accumulators of stage 0 (columns 0..255)
        const uint32_t base = tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 32u;
        uint32_t acc = 0;
        const long long t0 = clock64();
        if (p.readers)
            for (uint32_t r = 0; r < p.rounds; ++r)
                for (uint32_t a = 0; a < 4; ++a)
                    for (uint32_t c = 0; c < 32; c += W) acc += ld_cols<W>(base + a * 64u + c);
        const long long t1 = clock64();
        if (acc == 0x12345u) p.cycles[0] = 1;
        if (lane == 0 && warp == 0) p.cycles[2 * blockIdx.x] = (unsigned long long) (t1 - t0);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) stop_s = 1;
    } else if (lane == 0) {
        // MMA warp: N = 128 MMAs into stage 1 (columns 256..511) until the readers are done (or `rounds` x 28 when there are none)
        const uint64_t da = make_desc(smem_u32(smem), 1152, 144), db = make_desc(smem_u32(smem + 32768), 2048, 128);
        const uint32_t idesc = make_idesc(128, 128);
        if (p.mma) {
            const unsigned long long limit = p.readers ? ~0ull : (unsigned long long) p.rounds * 28ull;
            const long long t0 = clock64();
            while (n_mma < limit && (p.readers == 0 || stop_s == 0)) {
                for (uint32_t j = 0; j < 4; ++j)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, q;\n\t}\n"
                                 ::"r"(tmem + 256u + (j & 1u) * 64u), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                n_mma += 4;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (!p.readers) p.cycles[2 * blockIdx.x] = (unsigned long long) (clock64() - t0);
        }
        p.cycles[2 * blockIdx.x + 1] = n_mma;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int smem = 64 * 1024, grid = prop.multiProcessorCount;
    CK(cudaFuncSetAttribute(rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(rate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(rate_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long *d_cyc, h[1024];
    CK(cudaMalloc(&d_cyc, sizeof(h)));
    const uint32_t rounds = 400;
    for (uint32_t mma : {0u, 1u})
        for (uint32_t width : {8u, 16u, 32u}) {
            Args a{width, rounds, mma, 1u, d_cyc};
            CK(cudaMemset(d_cyc, 0, sizeof(h)));
            if (width == 8) rate_kernel<8><<<grid, 288, smem>>>(a);
            else if (width == 16) rate_kernel<16><<<grid, 288, smem>>>(a);
            else rate_kernel<32><<<grid, 288, smem>>>(a);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_cyc, 16 * grid, cudaMemcpyDeviceToHost));
            unsigned long long mx = 0, mm = 0;
            for (int i = 0; i < grid; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; mm += h[2 * i + 1]; }
            const double per_tile = (double) mx / rounds;      // cycles for the 8 warps to read 128 KB (one tile)
            printf("ld width x%-2u %s: %7.0f cycles per tile (128 KB, 8 warps) = %5.1f B/clk/SM", width, mma ? "with MMAs   " : "no MMAs     ", per_tile, 131072.0 / per_tile);
            if (mma) printf(";  MMAs meanwhile: %.1f cycles per N=128 MMA", (double) mx / ((double) mm / grid));
            printf("\n");
        }
    {
        Args a{8u, rounds, 1u, 0u, d_cyc};
        rate_kernel<8><<<grid, 288, smem>>>(a);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_cyc, 16 * grid, cudaMemcpyDeviceToHost));
        printf("MMAs alone: %.1f cycles per N=128 MMA\n", (double) h[0] / (double) h[1]);
    }
    return 0;
}
