// Cost (SM cycles seen by the issuing thread) of the synchronisation instructions the ring kernel's MMA warp executes per tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/instr_cost tools/instr_cost.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mbar_test(uint32_t a, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done;
}
__global__ void __launch_bounds__(128, 1) cost_kernel(long long *out, int busy_warps) {
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int R = 200;
    if (threadIdx.x == 0) {
        long long t0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t sink = 0;
        for (int r = 0; r < R; ++r) {
            t0 = clock64(); acc[0] += clock64() - t0;                                                   // clock64 itself
            t0 = clock64(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); acc[1] += clock64() - t0;
            t0 = clock64(); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); acc[2] += clock64() - t0;
            t0 = clock64(); asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory"); acc[3] += clock64() - t0;
            t0 = clock64(); sink += mbar_test(smem_u32(&bar2), 0); acc[4] += clock64() - t0;              // unsatisfied test
            t0 = clock64(); __nanosleep(40); acc[5] += clock64() - t0;
            t0 = clock64(); while (!mbar_test(smem_u32(&bar), r & 1)) {} acc[6] += clock64() - t0;        // wait for the commit above
            t0 = clock64(); asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory"); sink += mbar_test(smem_u32(&bar2), r & 1); acc[7] += clock64() - t0;
        }
        for (int i = 0; i < 8; ++i) out[i] = acc[i] / R;
        out[8] = sink;
        stop = 1;
    } else if ((int) (threadIdx.x >> 5) <= busy_warps && (threadIdx.x >> 5) >= 1) {
        // optional: other warps spin on shared memory next to the measured thread (like pollers in the ring kernel)
        while (!stop) __nanosleep(40);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}
int main() {
    long long *d, h[9];
    CK(cudaMalloc(&d, 72));
    const char *names[] = {"clock64 pair", "tcgen05.fence::after_thread_sync", "tcgen05.fence::before_thread_sync", "tcgen05.commit (issue)", "mbarrier.test_wait (not done)",
                           "nanosleep(40)", "spin until own commit arrives", "mbarrier.arrive + test"};
    for (int busy : {0, 3}) {
        cost_kernel<<<1, 128>>>(d, busy);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost));
        printf("-- %d other warps polling with nanosleep\n", busy);
        for (int i = 0; i < 8; ++i) printf("%-40s %6lld cycles\n", names[i], h[i]);
    }
    return 0;
}
