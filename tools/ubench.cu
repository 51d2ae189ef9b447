// Micro-benchmarks that calibrate the roofline of the cloud kernel on the actual part:
//   1. IMAD issue rate (the K1 kernel does 3n+1 dependent-free IMADs per output word)
//   2. achievable HBM bandwidth for K1's traffic shape: read 1 byte per ~5 bytes written, streaming
//      128-bit stores (st.global.cs), versus a plain copy
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void imad_kernel(uint32_t *out, uint32_t c, int iters) {
    uint32_t acc[ILP];
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * c + x;   // IMAD
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= acc[i];
    if (s == 0x12345678u) out[0] = s;
}

// mixed: IMAD (fma pipe) + IADD3/LOP3 (alu pipe) to see whether they dual-issue
template <int ILP>
__global__ void imad_iadd_kernel(uint32_t *out, uint32_t c, int iters) {
    uint32_t acc[ILP], acc2[ILP];
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = x + i; acc2[i] = x ^ i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { acc[i] = acc[i] * c + x; acc2[i] = (acc2[i] ^ x) + acc2[i]; }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= acc[i] ^ acc2[i];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void copy_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __ldg(in + i);
}

// K1's shape: each input vector is expanded into `fan` output vectors (write-dominated)
__global__ void fanout_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n_in, int fan) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n_in; i += stride) {
        const uint4 v = __ldg(in + i);
        for (int f = 0; f < fan; ++f) {
            uint4 w = make_uint4(v.x + f, v.y, v.z, v.w);
            asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t) f * n_in + i), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
        }
    }
}

__global__ void write_kernel(uint4 *__restrict__ out, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + i), "r"((uint32_t) i), "r"(1u), "r"(2u), "r"(3u) : "memory");
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device: %s, %d SMs, max clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint32_t *d_out;
    CK(cudaMalloc(&d_out, 1024));
    float ms;

    const int iters = 4096;
    const int grid = prop.multiProcessorCount * 8, block = 256;
    for (int rep = 0; rep < 2; ++rep) {
        imad_kernel<8><<<grid, block>>>(d_out, 3, iters);
        CK(cudaEventRecord(e0));
        imad_kernel<8><<<grid, block>>>(d_out, 3, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    double ops = (double) grid * block * iters * 8;
    printf("IMAD: %.2f T IMAD/s  (%.1f per clk per SM at %d MHz nominal)\n", ops / ms * 1e-9,
           ops / (ms * 1e-3) / prop.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000);
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        imad_iadd_kernel<8><<<grid, block>>>(d_out, 3, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("IMAD + 2 ALU ops interleaved: %.2f T IMAD/s (if ~equal to the line above, fma and alu pipes overlap)\n", ops / ms * 1e-9);

    const size_t n_in = (size_t) 400 << 20 >> 4;       // 400 MiB of uint4
    const int fan = 5;
    uint4 *d_in, *d_big;
    CK(cudaMalloc(&d_in, n_in * 16));
    CK(cudaMalloc(&d_big, n_in * 16 * fan));
    CK(cudaMemset(d_in, 1, n_in * 16));
    const int g2 = prop.multiProcessorCount * 16;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        copy_kernel<<<g2, 256>>>(d_in, d_big, n_in);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("copy 400 MiB -> 400 MiB: %.0f GB/s (read+write)\n", 2.0 * n_in * 16 / ms * 1e-6);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        fanout_kernel<<<g2, 256>>>(d_in, d_big, n_in, fan);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("fan-out 1 read : %d writes (K1 shape), %.2f GB total: %.0f GB/s, %.3f ms\n", fan, (1.0 + fan) * n_in * 16 * 1e-9,
           (1.0 + fan) * n_in * 16 / ms * 1e-6, ms);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        write_kernel<<<g2, 256>>>(d_big, n_in * fan);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("write-only 2 GB: %.0f GB/s\n", (double) fan * n_in * 16 / ms * 1e-6);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(d_big, d_big + n_in * 2, n_in * 16 * 2, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("cudaMemcpy D2D 800 MiB: %.0f GB/s (read+write)\n", 2.0 * n_in * 16 * 2 / ms * 1e-6);
    // PCIe
    void *h;
    CK(cudaMallocHost(&h, n_in * 16));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(d_in, h, n_in * 16, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("H2D pinned 400 MiB: %.1f GB/s\n", n_in * 16 / ms * 1e-6);
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(h, d_in, n_in * 16, cudaMemcpyDeviceToHost));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("D2H pinned 400 MiB: %.1f GB/s\n", n_in * 16 / ms * 1e-6);
    return 0;
}
