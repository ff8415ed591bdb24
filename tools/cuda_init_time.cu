// Baseline cost of a CUDA process on this box: cuInit + primary context (cudaFree(0)), then 1 GiB cudaMalloc + memset.
// The pastar CLI pays the first inside "Phase 1" (tools/cli_kinase_time.py); nothing in the library can shorten it.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    auto t0 = std::chrono::high_resolution_clock::now();
    cudaFree(0);
    auto t1 = std::chrono::high_resolution_clock::now();
    void *p = nullptr;
    cudaMalloc(&p, 1ull << 30);
    cudaMemset(p, 0, 1ull << 30);
    cudaDeviceSynchronize();
    auto t2 = std::chrono::high_resolution_clock::now();
    printf("cuda init + primary context %.3f s; 1 GiB malloc + memset %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
           std::chrono::duration<double>(t2 - t1).count());
    return 0;
}
