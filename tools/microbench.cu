// Measured roofline for the dedupe probe: random 16-byte gathers / atomics over a table much larger than L2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x *= 0x9E3779B97F4A7C15ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; return x;
}
__device__ __forceinline__ void ld16(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

template <int MLP>
__global__ void gather_kernel(const unsigned long long *tab, unsigned long long mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long acc = 0, seed = tid * 0x1234567ull + 1;
    for (int it = 0; it < iters; it++) {
        unsigned long long a[MLP], b[MLP];
#pragma unroll
        for (int j = 0; j < MLP; j++) {
            seed = mix(seed + j);
            ld16(tab + 2 * (seed & mask), a[j], b[j]);
        }
#pragma unroll
        for (int j = 0; j < MLP; j++) acc += a[j] ^ b[j];
    }
    if (acc == 0x1234) out[0] = acc;
}

__global__ void atomic_or_kernel(unsigned long long *tab, unsigned long long mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long acc = 0, seed = tid * 0x1234567ull + 1;
    for (int it = 0; it < iters; it++) {
        seed = mix(seed);
        acc += atomicOr(tab + 2 * (seed & mask) + 1, 1ull << 31);
    }
    if (acc == 0x1234) out[0] = acc;
}

__global__ void hot_atomic_kernel(unsigned long long *ctr, int nctr, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    for (int it = 0; it < iters; it++) acc += atomicAdd(ctr + ((tid + it) % nctr) * 16, 1ull);  // 128 B apart
    if (acc == 0x1234) out[0] = acc;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main(int argc, char **argv)
{
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu\n", lim);
    const unsigned long long entries = 1ull << 30; // 16 GiB
    unsigned long long *tab, *out, *ctr;
    CK(cudaMalloc(&tab, entries * 16)); CK(cudaMemset(tab, 0, entries * 16));
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&ctr, 1 << 20)); CK(cudaMemset(ctr, 0, 1 << 20));
    for (int lg = 22; lg <= 30; lg += 2) { // table size sweep: 64 MiB .. 16 GiB
        int blocks = 148 * 4, threads = 256, iters = 64;
        float ms = timeit([&] { gather_kernel<8><<<blocks, threads>>>(tab, (1ull << lg) - 1, iters, out); });
        double n = (double)blocks * threads * iters * 8;
        printf("gather16 table %6.0f MiB MLP=8: %.2f G loads/s\n", (double)(16ull << lg) / (1 << 20), n / ms / 1e6);
    }
    for (int occ : {4}) {
        int blocks = 148 * occ, threads = 256, iters = 64;
        double n;
        float ms;
        ms = timeit([&] { gather_kernel<1><<<blocks, threads>>>(tab, entries - 1, iters * 8, out); }); n = (double)blocks * threads * iters * 8;
        printf("gather16 occ=%d CTAs/SM MLP=1: %.3f ms  %.2f G loads/s  (%.0f GB/s at 32 B/sector)\n", occ, ms, n / ms / 1e6, n * 32 / ms / 1e6);
        ms = timeit([&] { gather_kernel<4><<<blocks, threads>>>(tab, entries - 1, iters * 2, out); }); n = (double)blocks * threads * iters * 8;
        printf("gather16 occ=%d CTAs/SM MLP=4: %.3f ms  %.2f G loads/s  (%.0f GB/s)\n", occ, ms, n / ms / 1e6, n * 32 / ms / 1e6);
        ms = timeit([&] { gather_kernel<8><<<blocks, threads>>>(tab, entries - 1, iters, out); }); n = (double)blocks * threads * iters * 8;
        printf("gather16 occ=%d CTAs/SM MLP=8: %.3f ms  %.2f G loads/s  (%.0f GB/s)\n", occ, ms, n / ms / 1e6, n * 32 / ms / 1e6);
        ms = timeit([&] { gather_kernel<16><<<blocks, threads>>>(tab, entries - 1, iters / 2, out); }); n = (double)blocks * threads * iters * 8;
        printf("gather16 occ=%d CTAs/SM MLP=16: %.3f ms  %.2f G loads/s  (%.0f GB/s)\n", occ, ms, n / ms / 1e6, n * 32 / ms / 1e6);
        ms = timeit([&] { atomic_or_kernel<<<blocks, threads>>>(tab, entries - 1, iters, out); }); n = (double)blocks * threads * iters;
        printf("atomicOr random occ=%d: %.3f ms  %.2f G atomics/s\n", occ, ms, n / ms / 1e6);
    }
    for (int nctr : {1, 8, 64, 1024}) {
        int blocks = 148 * 4, threads = 256, iters = 64;
        float ms = timeit([&] { hot_atomic_kernel<<<blocks, threads>>>(ctr, nctr, iters, out); });
        double n = (double)blocks * threads * iters;
        printf("atomicAdd(return) on %d hot counters: %.3f ms  %.2f G atomics/s\n", nctr, ms, n / ms / 1e6);
    }
    return 0;
}
