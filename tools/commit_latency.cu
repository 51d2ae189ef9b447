// How long after tcgen05.commit does ANOTHER warp see the mbarrier complete, depending on what the issuing thread does
// next? (The ring kernel's trace shows t_full arriving ~2500 cycles after the commit was issued.)
//   mode 0: issuer busy-loops on clock64            mode 1: issuer __nanosleep(40) loop
//   mode 2: issuer polls an unrelated mbarrier with test_wait + nanosleep(40)  (what the MMA warp does)
//   mode 3: issuer polls with try_wait (hardware-suspended)                    mode 4: issuer waits on the same barrier itself
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/commit_latency tools/commit_latency.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t) 1 << 46);
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t a, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t a, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done;
}

struct Args { uint32_t mode, nm, rounds, waiter_try; long long *lat; };

__global__ void __launch_bounds__(128, 1) lat_kernel(const Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, other, back;
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_commit;
    for (uint32_t i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&other)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&back)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t warp = threadIdx.x >> 5;
    long long sum = 0, mx = 0;
    if (warp == 0) {
        if (threadIdx.x == 0) {
            const uint64_t da = make_desc(smem_u32(smem), 1152, 144), db = make_desc(smem_u32(smem + 32768), 2048, 128);
            const uint32_t idesc = (2u << 4) | (1u << 10) | (1u << 15) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            for (uint32_t r = 0; r < p.rounds; ++r) {
                for (uint32_t i = 0; i < p.nm; ++i)
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, q;\n\t}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i) : "memory");
                *(volatile long long *) &t_commit = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                // what the issuer does while the commit is in flight; it leaves when the waiter hands `back` over
                const uint32_t ba = smem_u32(&back), par = r & 1u;
                if (p.mode == 0) { while (!mbar_test(ba, par)) { const long long t = clock64(); while (clock64() - t < 40) {} } }
                else if (p.mode == 1) { while (!mbar_test(ba, par)) __nanosleep(40); }
                else if (p.mode == 2) { while (!mbar_test(ba, par)) { (void) mbar_test(smem_u32(&other), 0); __nanosleep(40); } }
                else if (p.mode == 3) { while (!mbar_try(ba, par)) {} }
                else { while (!mbar_try(smem_u32(&bar), par)) {} while (!mbar_try(ba, par)) {} }
            }
        }
    } else if (warp == 1) {
        if ((threadIdx.x & 31) == 0) {
            const uint32_t a = smem_u32(&bar);
            for (uint32_t r = 0; r < p.rounds; ++r) {
                if (p.waiter_try) { while (!mbar_try(a, r & 1u)) {} } else { while (!mbar_test(a, r & 1u)) __nanosleep(40); }
                const long long d = clock64() - *(volatile long long *) &t_commit;
                if (r >= 8) { sum += d; mx = d > mx ? d : mx; }
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&back)) : "memory");
            }
            p.lat[blockIdx.x * 2] = sum / (p.rounds - 8);
            p.lat[blockIdx.x * 2 + 1] = mx;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

int main() {
    CK(cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    long long *d, h[2];
    CK(cudaMalloc(&d, 16));
    const char *names[] = {"issuer busy-loops", "issuer nanosleep(40) loop", "issuer polls other barrier + nanosleep", "issuer try_wait on hand-back", "issuer waits on the same barrier"};
    for (uint32_t waiter_try : {0u, 1u})
        for (uint32_t nm : {0u, 8u})
            for (uint32_t mode = 0; mode < 5; ++mode) {
                Args a{mode, nm, 200, waiter_try, d};
                lat_kernel<<<1, 128, 64 * 1024>>>(a);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                printf("waiter %-9s nm=%u  %-40s commit -> seen by another warp: mean %5lld  max %5lld cycles\n", waiter_try ? "try_wait" : "test+sleep", nm, names[mode], h[0], h[1]);
            }
    return 0;
}
