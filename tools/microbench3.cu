// Round-2 questions about the closed/open-table probe ceiling (tools/microbench2.cu: 36.5 G random 128-byte lines/s):
//   A. does the line rate rise when the lines a thread reads are ADJACENT (same 256 B .. 4 KB aligned block)?
//      -> decides whether blocking the table at a coarser granularity than one line pays
//   B. 4-byte loads instead of 16-byte ones (same lines)
//   C. a dependent pair: 8-byte header from an L2-sized array, then the data line the header names
// Each thread keeps 8 independent loads in flight, as the expand kernel does.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{ x *= 0x9E3779B97F4A7C15ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; return x; }
__device__ __forceinline__ void ld16(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{ asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ unsigned ld4(const unsigned *p)
{ unsigned v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld8(const unsigned long long *p)
{ unsigned long long v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// A: blocks of BL lines; a thread takes U consecutive lines (wrapping) of 8/U random blocks per iteration
template <int BL, int U>
__global__ void kA(const unsigned long long *tab, unsigned long long nblocks_mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long acc = 0, seed = tid * 0x1234567ull + 1;
    for (int it = 0; it < iters; it++) {
        unsigned long long a[8], b[8];
#pragma unroll
        for (int q = 0; q < 8 / U; q++) {
            seed = mix(seed + q);
            const unsigned long long blk = seed & nblocks_mask;
            const unsigned start = (unsigned)(seed >> 40) & (BL - 1);
            const unsigned ent = (unsigned)(seed >> 50) & 7u;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned long long line = blk * BL + ((start + u) & (BL - 1));
                ld16(tab + line * 16 + ent * 2, a[q * U + u], b[q * U + u]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += a[j] ^ b[j];
    }
    if (acc == 0x1234) out[0] = acc;
}
template <int BL, int U> void runA(const unsigned long long *tab, unsigned long long lines, unsigned long long *out)
{
    int blocks = 148 * 4, threads = 256, iters = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kA<BL, U><<<blocks, threads>>>(tab, lines / BL - 1, iters, out); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a); kA<BL, U><<<blocks, threads>>>(tab, lines / BL - 1, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double loads = (double)blocks * threads * iters * 8;
    printf("A block %5d B, %d adjacent lines used per block : %.2f G lines/s\n", BL * 128, U, loads / best / 1e6);
}

// B: 4-byte loads, random lines
__global__ void kB(const unsigned *tab, unsigned long long mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long seed = tid * 0x1234567ull + 1;
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { seed = mix(seed + j); a[j] = ld4(tab + (seed & mask) * 32 + ((seed >> 45) & 31)); }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += a[j];
    }
    if (acc == 0x1234) out[0] = acc;
}
// C: header (8 B, from an array of hbytes) -> dependent data line (4 B load) ; the header holds the line index
__global__ void kC(const unsigned long long *hdr, unsigned long long hmask, const unsigned *tab, unsigned long long mask, int iters, unsigned long long *out)
{
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long seed = tid * 0x1234567ull + 1;
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned long long h[8];
        unsigned a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { seed = mix(seed + j); h[j] = ld8(hdr + (seed & hmask)); }
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = ld4(tab + (h[j] & mask) * 32 + ((h[j] >> 45) & 31));
#pragma unroll
        for (int j = 0; j < 8; j++) acc += a[j];
    }
    if (acc == 0x1234) out[0] = acc;
}
__global__ void fill_hdr(unsigned long long *hdr, unsigned long long n)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) hdr[i] = mix(i * 77 + 5);
}
int main()
{
    const unsigned long long bytes = 16ull << 30, lines = bytes / 128;
    unsigned long long *tab, *out;
    if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(tab, 0, bytes); cudaMalloc(&out, 64);
    runA<1, 1>(tab, lines, out);
    runA<2, 2>(tab, lines, out); runA<4, 2>(tab, lines, out); runA<4, 4>(tab, lines, out);
    runA<8, 2>(tab, lines, out); runA<8, 4>(tab, lines, out); runA<8, 8>(tab, lines, out);
    runA<16, 2>(tab, lines, out); runA<16, 4>(tab, lines, out); runA<16, 8>(tab, lines, out);
    runA<32, 4>(tab, lines, out); runA<32, 8>(tab, lines, out);
    runA<64, 8>(tab, lines, out); runA<256, 8>(tab, lines, out);
    {
        int blocks = 148 * 4, threads = 256, iters = 64;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int tsz = 0; tsz < 2; tsz++) { // 16 GiB and 4 GiB tables
            const unsigned long long m = (tsz == 0 ? lines : lines / 4) - 1;
            kB<<<blocks, threads>>>((const unsigned *)tab, m, iters, out); cudaDeviceSynchronize();
            cudaEventRecord(a); kB<<<blocks, threads>>>((const unsigned *)tab, m, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("B 4-byte loads, random lines of a %d GiB table : %.2f G lines/s\n", tsz == 0 ? 16 : 4, (double)blocks * threads * iters * 8 / ms / 1e6);
        }
        unsigned long long *hdr; cudaMalloc(&hdr, 1ull << 30);
        for (int hsz = 0; hsz < 4; hsz++) {
            const unsigned long long hb = (32ull << 20) << (hsz * 1); // 32, 64, 128, 256 MiB of headers
            fill_hdr<<<592, 256>>>(hdr, hb / 8);
            for (int tsz = 0; tsz < 2; tsz++) {
                const unsigned long long m = (tsz == 0 ? lines : lines / 4) - 1;
                kC<<<blocks, threads>>>(hdr, hb / 8 - 1, (const unsigned *)tab, m, iters, out); cudaDeviceSynchronize();
                cudaEventRecord(a); kC<<<blocks, threads>>>(hdr, hb / 8 - 1, (const unsigned *)tab, m, iters, out); cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                printf("C header array %4llu MiB (random 8 B) -> dependent random data line, %d GiB table : %.2f G pairs/s\n", hb >> 20, tsz == 0 ? 16 : 4,
                       (double)blocks * threads * iters * 8 / ms / 1e6);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
