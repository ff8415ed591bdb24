// How does the random-access ceiling depend on how much of each 128-byte line is used?
// Each group of G consecutive lanes reads G consecutive 16-byte entries of one random 128-byte-aligned line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{ x *= 0x9E3779B97F4A7C15ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; return x; }
__device__ __forceinline__ void ld16(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{ asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
template <int G>
__global__ void k(const unsigned long long *tab, unsigned long long mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long grp = tid / G, sub = tid % G;
    unsigned long long acc = 0, seed = grp * 0x1234567ull + 1;
    for (int it = 0; it < iters; it++) {
        unsigned long long a[8], b[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            seed = mix(seed + j);
            unsigned long long line = seed & mask;                 // 128-byte line index
            ld16(tab + line * 16 + sub * 2, a[j], b[j]);           // 16 u64 per line; entry `sub` of the line
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += a[j] ^ b[j];
    }
    if (acc == 0x1234) out[0] = acc;
}
template <int G> void run(const unsigned long long *tab, unsigned long long lines, unsigned long long *out)
{
    int blocks = 148 * 4, threads = 256, iters = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<G><<<blocks, threads>>>(tab, lines - 1, iters, out); cudaDeviceSynchronize();
    cudaEventRecord(a); k<G><<<blocks, threads>>>(tab, lines - 1, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double loads = (double)blocks * threads * iters * 8;
    printf("entries used per random line = %d : %.2f G entry-loads/s = %.2f G lines/s (%.0f GB/s of entries)\n", G, loads / ms / 1e6, loads / G / ms / 1e6, loads * 16 / ms / 1e6);
}
int main()
{
    const unsigned long long bytes = 16ull << 30, lines = bytes / 128;
    unsigned long long *tab, *out;
    cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes); cudaMalloc(&out, 64);
    run<1>(tab, lines, out); run<2>(tab, lines, out); run<4>(tab, lines, out); run<8>(tab, lines, out);
    return 0;
}
