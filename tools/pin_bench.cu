// pin_bench -- what the host staging of 2 GB of output ciphertexts costs on this box: CUDA context creation, cudaHostAlloc,
// pre-faulted pageable memory + cudaHostRegister, and the D2H copy into pinned / registered / pageable memory.
// nvcc -O2 -o tools/pin_bench tools/pin_bench.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <sys/mman.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void touch(char *p, size_t n, int nt) {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=]() { for (size_t i = n * t / nt; i < n * (t + 1) / nt; i += 4096) p[i] = 0; });
    for (auto &x : th) x.join();
}
int main() {
    const size_t n = (size_t) 1991638376;
    const int nt = (int) std::thread::hardware_concurrency();
    double t = now();
    cudaFree(0);
    printf("context creation            %.3f s\n", now() - t);
    char *d; cudaMalloc(&d, n); cudaMemset(d, 1, n); cudaDeviceSynchronize();
    t = now(); char *h1; cudaHostAlloc(&h1, n, cudaHostAllocDefault); printf("cudaHostAlloc 2 GB          %.3f s\n", now() - t);
    t = now(); cudaMemcpy(h1, d, n, cudaMemcpyDeviceToHost); printf("D2H to cudaHostAlloc        %.3f s\n", now() - t);
    t = now(); cudaFreeHost(h1); printf("cudaFreeHost                %.3f s\n", now() - t);
    t = now(); char *h2 = (char *) aligned_alloc(4096, (n + 4095) & ~(size_t) 4095); touch(h2, n, nt); printf("malloc + touch (%d thr)     %.3f s\n", nt, now() - t);
    t = now(); cudaMemcpy(h2, d, n, cudaMemcpyDeviceToHost); printf("D2H to pageable (touched)   %.3f s\n", now() - t);
    t = now(); cudaError_t e = cudaHostRegister(h2, n, cudaHostRegisterDefault); printf("cudaHostRegister            %.3f s (%s)\n", now() - t, cudaGetErrorString(e));
    t = now(); cudaMemcpy(h2, d, n, cudaMemcpyDeviceToHost); printf("D2H to registered           %.3f s\n", now() - t);
    cudaHostUnregister(h2);
    t = now();
    char *h3 = (char *) mmap(nullptr, (n + (2 << 20)) & ~(size_t) ((2 << 20) - 1), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(h3, n, MADV_HUGEPAGE); touch(h3, n, nt); printf("mmap + THP + touch          %.3f s\n", now() - t);
    t = now(); e = cudaHostRegister(h3, n, cudaHostRegisterDefault); printf("cudaHostRegister (THP)      %.3f s (%s)\n", now() - t, cudaGetErrorString(e));
    t = now(); cudaMemcpy(h3, d, n, cudaMemcpyDeviceToHost); printf("D2H to registered THP       %.3f s\n", now() - t);
    t = now(); char *h4 = (char *) aligned_alloc(4096, (n + 4095) & ~(size_t) 4095); cudaMemcpy(h4, d, n, cudaMemcpyDeviceToHost); printf("malloc + D2H to untouched    %.3f s\n", now() - t);
    return 0;
}
