// Kernel 1 — all N(N-1)/2 reverse pairwise DP cost tables in one launch.
//
// Replaces the serial double loop of PairAlign::Align / pairCost / gapCost
// (reference pastar/PairAlign.cpp:137-171, 107-134, 96-101) called once per
// pair from HeuristicHPair::init (pastar/HeuristicHPair.cpp:54-61).
//
//   M[i][j] = min( M[i+1][j] + gap(dir[i+1][j], X),      "c0"
//                  M[i][j+1] + gap(dir[i][j+1], Y),      "c1"  (c0 < c1 ? X : Y)
//                  M[i+1][j+1] + cost(s1[i], s2[j]) )    "c2"  (strictly smaller wins)
//   borders M[L1][j], M[i][L2] = GapOpen + k * GapExtension, M[L1][L2] = 0.
//
// The 2-bit direction of the winning move is carried next to every value so
// the single-matrix affine rule stays general even though the reference's
// constants make it numerically dead (Cost.h:13: open == extension == 30).
//
// Mapping: one CTA per pair.  A warp owns a band of 32 consecutive rows, one
// row per lane, and sweeps it right-to-left as an anti-diagonal wavefront:
// at step s lane l computes column L2-1-(s-l).  The cell below comes from lane
// l-1 by __shfl_up (value and direction packed in one register); the band
// below (another warp of the CTA) streams its top row through a shared-memory
// ring with producer/consumer counters, so warps run skewed against each other
// with no CTA-wide barrier.  Results are staged in a 32x32 shared-memory tile
// per warp and written out as row segments of 32 cells.
//
// Bound: integer ALU / dependency latency (min-plus; no tensor-core shape).
// Bytes: one cell written once (4 B, or 2 B when 30*(L1+L2) < 65536).
#include "pg_internal.cuh"

namespace {

constexpr int DP_RING = 256;  // ring entries per warp (power of two)
constexpr int DP_CHUNK = 8;   // consumer fetch / producer publish granularity (steps)

enum { NoGap = 0, GapX = 1, GapY = 2 };

struct DpShared {
    int32_t cost[90 * 90];
};

template <typename TC>
__global__ void __launch_bounds__(1024, 1) pair_dp_kernel(const __grid_constant__ DevProblem p, int warps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *s_cost = reinterpret_cast<int32_t *>(smem_raw);
    volatile int32_t *s_ring = reinterpret_cast<volatile int32_t *>(s_cost + 90 * 90);
    int32_t *s_tile = const_cast<int32_t *>(s_ring) + warps * DP_RING;
    volatile unsigned *s_prod = reinterpret_cast<volatile unsigned *>(s_tile + warps * 32 * 33);
    volatile unsigned *s_cons = s_prod + warps;

    const int pair = blockIdx.x;
    const int sa = p.pa[pair], sb = p.pb[pair];
    const int L1 = p.len[sa], L2 = p.len[sb];
    const int cols = L2 + 1;
    TC *M = const_cast<TC *>(reinterpret_cast<const TC *>(p.table[pair]));
    const uint8_t *s1 = p.seq[sa];
    const uint8_t *s2 = p.seq[sb];
    const int open = p.gap_open, ext = p.gap_ext;

    for (int i = threadIdx.x; i < 90 * 90; i += blockDim.x) s_cost[i] = p.cost[i];
    if (threadIdx.x < warps) {
        s_prod[threadIdx.x] = 0;
        s_cons[threadIdx.x] = 0;
    }
    // borders, PairAlign.cpp:142-160
    for (int j = threadIdx.x; j <= L2; j += blockDim.x) M[(size_t)L1 * cols + j] = (TC)(j == L2 ? 0 : open + (L2 - 1 - j) * ext);
    for (int i = threadIdx.x; i < L1; i += blockDim.x) M[(size_t)i * cols + L2] = (TC)(open + (L1 - 1 - i) * ext);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbands = (L1 + 31) >> 5;
    if (L2 == 0) return;
    int32_t *tile = s_tile + warp * 32 * 33;
    volatile int32_t *ring_out = s_ring + warp * DP_RING;
    const int pwarp = (warp + warps - 1) % warps;
    volatile int32_t *ring_in = s_ring + pwarp * DP_RING;

    for (int band = warp; band < nbands; band += warps) {
        const int r = L1 - 1 - (band * 32 + lane); // my row; lane 0 is the bottom row of the band
        const bool active = r >= 0;
        const int32_t *costrow = s_cost + (active ? (int)s1[r] : 0) * 90;
        int right_v = open + (L1 - 1 - r) * ext; // M[r][L2]
        int right_d = GapX;
        int diag_v = (r + 1 == L1) ? 0 : open + (L1 - 2 - r) * ext; // M[r+1][L2]
        int my_pk = 0;
        const unsigned in_base = band > 0 ? (unsigned)((band - 1) / warps) * (unsigned)L2 : 0u;
        const unsigned out_base = (unsigned)(band / warps) * (unsigned)L2;
        const bool has_consumer = band + 1 < nbands;
        const int nsteps = L2 + 31;
        int nb = 0;
        int last_flush = -1;
        int cnext = active ? costrow[s2[L2 - 1]] : 0;

        for (int s = 0; s < nsteps; s++) {
            if ((s & (DP_CHUNK - 1)) == 0) {
                // --- fetch the next DP_CHUNK cells of the row below the band (for lane 0)
                const int kk = s + lane;
                if (band == 0) {
                    nb = ((open + kk * ext) << 2) | GapY; // M[L1][L2-1-kk], PairAlign.cpp:146-154
                } else if (s < L2) {
                    const unsigned need = in_base + (unsigned)min(s + DP_CHUNK, L2);
                    while ((int)(s_prod[pwarp] - need) < 0) { }
                    __threadfence_block();
                    if (lane < DP_CHUNK && kk < L2) nb = ring_in[(in_base + kk) & (DP_RING - 1)];
                    __syncwarp();
                    if (lane == 0) s_cons[pwarp] = need;
                }
                // --- back-pressure: the next DP_CHUNK values of lane 31 must fit in the ring
                if (has_consumer) {
                    const int k31 = s - 31;
                    if (k31 + DP_CHUNK > 0 && k31 < L2) {
                        const unsigned top = out_base + (unsigned)min(k31 + DP_CHUNK, L2);
                        while ((int)(top - s_cons[warp]) > DP_RING) { }
                    }
                }
            }
            const int k = s - lane;
            int down_pk = __shfl_up_sync(0xffffffffu, my_pk, 1);
            const int below = __shfl_sync(0xffffffffu, nb, s & (DP_CHUNK - 1));
            if (lane == 0) down_pk = below;
            if (active && k >= 0 && k < L2) {
                const int j = L2 - 1 - k;
                const int c = cnext;
                if (j > 0) cnext = costrow[s2[j - 1]];
                const int down_v = down_pk >> 2, down_d = down_pk & 3;
                const int c0 = down_v + (down_d == GapX ? ext : open);
                const int c1 = right_v + (right_d == GapY ? ext : open);
                int m, d;
                if (c0 < c1) {
                    m = c0;
                    d = GapX;
                } else {
                    m = c1;
                    d = GapY;
                }
                const int c2 = diag_v + c;
                if (c2 < m) {
                    m = c2;
                    d = NoGap;
                }
                diag_v = down_v;
                right_v = m;
                right_d = d;
                my_pk = (m << 2) | d;
                tile[lane * 33 + (k & 31)] = m;
                if (lane == 31 && has_consumer) ring_out[(out_base + k) & (DP_RING - 1)] = my_pk;
            }
            // --- publish the top row's progress
            if (has_consumer && ((s & (DP_CHUNK - 1)) == DP_CHUNK - 1 || s == nsteps - 1)) {
                __threadfence_block();
                const int k31 = s - 31;
                if (lane == 31 && k31 >= 0) s_prod[warp] = out_base + (unsigned)min(k31 + 1, L2);
            }
            // --- write the staged tile out as row segments
            if ((s & 31) == 31 || s == nsteps - 1) {
                __syncwarp();
                for (int rr = 0; rr < 32; rr++) {
                    const int row = L1 - 1 - (band * 32 + rr);
                    if (row < 0) break;
                    const int kk = last_flush + 1 - rr + lane;
                    if (kk >= 0 && kk <= s - rr && kk < L2) M[(size_t)row * cols + (L2 - 1 - kk)] = (TC)tile[rr * 33 + (kk & 31)];
                }
                last_flush = s;
                __syncwarp();
            }
        }
    }
}

} // namespace

int pg_launch_pair_dp(pg_ctx *ctx, float *kernel_ms)
{
    int max_bands = 1;
    for (const PairGeom &g : ctx->pairs) max_bands = std::max(max_bands, (g.rows - 1 + 31) / 32);
    const int warps = std::min(32, max_bands);
    const size_t smem = sizeof(int32_t) * 90 * 90 + (size_t)warps * DP_RING * 4 + (size_t)warps * 32 * 33 * 4 + (size_t)warps * 8;
    cudaEvent_t e0, e1;
    PG_CUDA(ctx, cudaEventCreate(&e0));
    PG_CUDA(ctx, cudaEventCreate(&e1));
    if (ctx->dp.cell16) {
        PG_CUDA(ctx, cudaFuncSetAttribute(pair_dp_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        PG_CUDA(ctx, cudaFuncSetAttribute(pair_dp_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    PG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (ctx->dp.cell16)
        pair_dp_kernel<uint16_t><<<ctx->npairs, warps * 32, smem, ctx->stream>>>(ctx->dp, warps);
    else
        pair_dp_kernel<int32_t><<<ctx->npairs, warps * 32, smem, ctx->stream>>>(ctx->dp, warps);
    PG_CUDA(ctx, cudaGetLastError());
    PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    PG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    return PG_OK;
}
