// Store-pattern micro-benchmark: what HBM write bandwidth can ONE persistent CTA per SM reach with the output
// shapes the cloud kernels use? The output is R rows (ciphertexts) of 8192 B; CTA (slice, chunk) owns a 512-byte
// column slice of the rows of its chunk and writes them 64 rows ("a tile") at a time.
//   P1  32-bit store per lane: a warp instruction writes 128 contiguous bytes of one row      (ring kernel today)
//   P2  128-bit store per lane: a warp instruction writes the whole 512-byte slice of one row
//   P3  cp.async.bulk shared -> global, one 512-byte copy per row, issued by one thread per tile (TMA store)
//   P4  like P3 but every CTA owns WHOLE rows (8192-byte bulk copies): the layout a row-major tiling would give
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_store tools/ubench_store.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) pattern_kernel(uint8_t *out, uint32_t n_rows, uint32_t n_chunks, uint32_t row_stride) {
    const uint32_t slice = blockIdx.x & 15u, chunk = blockIdx.x >> 4;
    const uint32_t r0 = (uint32_t) ((uint64_t) n_rows * chunk / n_chunks), r1 = (uint32_t) ((uint64_t) n_rows * (chunk + 1) / n_chunks);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *base = out + 512u * slice;
    if (MODE == 1) {
        // warp w: quadrant w & 3 (128 B of the slice), rows (w >> 2) * 32 .. +31 of every 64-row tile
        const uint32_t quad = warp & 3u, half = warp >> 2;
        for (uint32_t t = r0; t < r1; t += 64) {
#pragma unroll 8
            for (uint32_t i = 0; i < 32; ++i) {
                const uint32_t r = t + half * 32u + i;
                if (r < r1) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(base + (uint64_t) r * row_stride + quad * 128u + lane * 4u), "r"(r + lane) : "memory");
            }
        }
    } else if (MODE == 2) {
        for (uint32_t t = r0; t < r1; t += 64) {
#pragma unroll 8
            for (uint32_t i = 0; i < 8; ++i) {
                const uint32_t r = t + warp * 8u + i;
                if (r < r1) asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(base + (uint64_t) r * row_stride + lane * 16u), "r"(r), "r"(lane), "r"(2u), "r"(3u) : "memory");
            }
        }
    }
}

// P3 / P4: bulk copies from a shared-memory staging tile; ROW_BYTES per copy
template <uint32_t ROW_BYTES, uint32_t ROWS>
__global__ void __launch_bounds__(256, 1) bulk_kernel(uint8_t *out, uint32_t n_rows, uint32_t n_parts, uint32_t row_stride, uint32_t parts_per_row) {
    extern __shared__ __align__(128) uint8_t sm[];
    // two staging buffers of ROWS x ROW_BYTES
    const uint32_t col = blockIdx.x % parts_per_row, part = blockIdx.x / parts_per_row;
    const uint32_t r0 = (uint32_t) ((uint64_t) n_rows * part / n_parts), r1 = (uint32_t) ((uint64_t) n_rows * (part + 1) / n_parts);
    for (uint32_t i = threadIdx.x; i < 2 * ROWS * ROW_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t buf = 0;
    for (uint32_t t = r0; t < r1; t += ROWS) {
        // the warps would fill buffer `buf` here; emulate the hand-off
        __syncthreads();
        if (threadIdx.x < ROWS) {
            const uint32_t r = t + threadIdx.x;
            if (r < r1)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (uint64_t) r * row_stride + (uint64_t) col * ROW_BYTES),
                             "r"(smem_u32(sm + (buf * ROWS + threadIdx.x) * ROW_BYTES)), "r"(ROW_BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer written two tiles ago is free
        }
        buf ^= 1u;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const uint32_t n_rows = 242646, stride = 8192;
    uint8_t *d;
    CK(cudaMalloc(&d, (size_t) n_rows * stride));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    const uint32_t chunks = prop.multiProcessorCount / 16;
    const double gb = (double) n_rows * stride * 1e-9;
#define TIME(name, launch)                                                             \
    for (int rep = 0; rep < 4; ++rep) {                                                \
        CK(cudaEventRecord(e0));                                                       \
        launch;                                                                        \
        CK(cudaEventRecord(e1));                                                       \
        CK(cudaEventSynchronize(e1));                                                  \
        CK(cudaGetLastError());                                                        \
        CK(cudaEventElapsedTime(&ms, e0, e1));                                         \
    }                                                                                  \
    printf("%-70s %.3f ms  %.0f GB/s\n", name, ms, gb / (ms * 1e-3));
    TIME("P1 32-bit/lane stores, 16 slices x 9 chunks (144 CTAs)", (pattern_kernel<1><<<16 * chunks, 256>>>(d, n_rows, chunks, stride)));
    TIME("P2 128-bit/lane stores, 16 slices x 9 chunks", (pattern_kernel<2><<<16 * chunks, 256>>>(d, n_rows, chunks, stride)));
    CK(cudaFuncSetAttribute(bulk_kernel<512, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 512));
    TIME("P3 bulk 512 B x 64 rows per tile, 16 slices x 9 chunks", (bulk_kernel<512, 64><<<16 * chunks, 256, 2 * 64 * 512>>>(d, n_rows, chunks, stride, 16)));
    CK(cudaFuncSetAttribute(bulk_kernel<8192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 8192));
    TIME("P4 bulk 8192 B (whole rows) x 4 rows per tile, 148 CTAs", (bulk_kernel<8192, 4><<<prop.multiProcessorCount, 256, 2 * 4 * 8192>>>(d, n_rows, prop.multiProcessorCount, stride, 1)));
    CK(cudaFuncSetAttribute(bulk_kernel<2048, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 2048));
    TIME("P5 bulk 2048 B x 16 rows per tile, 4 slices x 37 chunks (148 CTAs)", (bulk_kernel<2048, 16><<<4 * (prop.multiProcessorCount / 4), 256, 2 * 16 * 2048>>>(d, n_rows, prop.multiProcessorCount / 4, stride, 4)));
    CK(cudaMemsetAsync(d, 1, (size_t) n_rows * stride));
    TIME("cudaMemset 2 GB", CK(cudaMemsetAsync(d, 1, (size_t) n_rows * stride)));
    return 0;
}
