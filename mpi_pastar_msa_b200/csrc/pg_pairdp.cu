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
#include <algorithm>
#include <cstdlib>

#include "pg_internal.cuh"

namespace {

constexpr int DP_RING = 256;  // ring entries per warp (power of two)
constexpr int DP_CHUNK = 8;   // consumer fetch / producer publish granularity (steps)
// rows per lane (a warp's band is 32 * DP_R rows) x warps per CTA (one 32*DP_R x 32 staging tile each): fewer rows per
// lane shorten the dependent chain of a step and put more warps on a pair; chosen by pg_launch_pair_dp

enum { NoGap = 0, GapX = 1, GapY = 2 };

// AFFINE = false is the reference's actual cost model (Cost.h:13: GapOpen == GapExtension): the direction state
// cannot change any value, so a cell is min3(below + gap, right + gap, diag + cost) - two dependent integer ops.
// AFFINE = true carries the direction of the winning move (PairAlign.cpp:96-134) for open != extension.
template <typename TC, bool AFFINE, int DP_R, int DP_MAXW>
__global__ void __launch_bounds__(32 * DP_MAXW, 1) pair_dp_kernel(const __grid_constant__ DevProblem p, int warps)
{
    constexpr int R = DP_R, BAND = 32 * DP_R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *s_cost = reinterpret_cast<int32_t *>(smem_raw);
    volatile int32_t *s_ring = reinterpret_cast<volatile int32_t *>(s_cost + 90 * 90);
    int32_t *s_tile = const_cast<int32_t *>(s_ring) + warps * DP_RING;
    volatile unsigned *s_prod = reinterpret_cast<volatile unsigned *>(s_tile + warps * BAND * 33);
    volatile unsigned *s_cons = s_prod + warps;

    const int pair = blockIdx.x;
    const int sa = p.pa[pair], sb = p.pb[pair];
    const int L1 = p.len[sa], L2 = p.len[sb];
    const int cols = L2 + 1;
    TC *M = const_cast<TC *>(reinterpret_cast<const TC *>(p.table[pair]));
    const uint8_t *s1 = p.seq[sa];
    const uint8_t *s2 = p.seq[sb];
    const int open = p.gap_open, ext = p.gap_ext;

    {
        const int4 *src = reinterpret_cast<const int4 *>(p.cost);
        int4 *dst = reinterpret_cast<int4 *>(s_cost);
        for (int i = threadIdx.x; i < 90 * 90 / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < warps) {
        s_prod[threadIdx.x] = 0;
        s_cons[threadIdx.x] = 0;
    }
    // borders, PairAlign.cpp:142-160
    for (int j = threadIdx.x; j <= L2; j += blockDim.x) M[(size_t)L1 * cols + j] = (TC)(j == L2 ? 0 : open + (L2 - 1 - j) * ext);
    for (int i = threadIdx.x; i < L1; i += blockDim.x) M[(size_t)i * cols + L2] = (TC)(open + (L1 - 1 - i) * ext);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbands = (L1 + BAND - 1) / BAND;
    if (L2 == 0) return;
    int32_t *tile = s_tile + warp * BAND * 33;
    volatile int32_t *ring_out = s_ring + warp * DP_RING;
    const int pwarp = (warp + warps - 1) % warps;
    volatile int32_t *ring_in = s_ring + pwarp * DP_RING;

    for (int band = warp; band < nbands; band += warps) {
        // lane l owns rows r0, r0-1, ..., r0-(R-1); k = 0 is the lowest of them, lane 0 / k 0 the bottom row of the band
        const int r0 = L1 - 1 - (band * BAND + lane * R);
        const int32_t *costrow[R];
        int right_v[R], right_d[R], cst[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int r = r0 - k;
            costrow[k] = s_cost + (r >= 0 ? (int)s1[r] : 0) * 90;
            right_v[k] = open + (L1 - 1 - r) * ext; // M[r][L2]
            right_d[k] = GapX;
            cst[k] = costrow[k][s2[L2 - 1]];
        }
        int diag0 = (r0 + 1 == L1) ? 0 : open + (L1 - 2 - r0) * ext; // M[r0+1][L2]
        int top_pk = 0;
        const unsigned in_base = band > 0 ? (unsigned)((band - 1) / warps) * (unsigned)L2 : 0u;
        const unsigned out_base = (unsigned)(band / warps) * (unsigned)L2;
        const bool has_consumer = band + 1 < nbands;
        const int nsteps = L2 + 31;
        const int rows_here = min(BAND, L1 - band * BAND); // rows of this band that exist
        int nb = 0;
        int last_flush = -1;
        int bnext = L2 >= 2 ? (int)s2[L2 - 2] : 0;

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
            const int kc = s - lane; // my column counter: column L2-1-kc
            int below_pk = __shfl_up_sync(0xffffffffu, top_pk, 1);
            const int below0 = __shfl_sync(0xffffffffu, nb, s & (DP_CHUNK - 1));
            if (lane == 0) below_pk = below0;
            if (kc >= 0 && kc < L2) {
                const int j = L2 - 1 - kc;
                const int bcur = bnext;               // residue of column j-1, loaded one step ahead
                bnext = j > 1 ? (int)s2[j - 2] : 0;
                int below_v = below_pk >> 2, below_d = below_pk & 3;
                int dg = diag0;
                diag0 = below_v;
                int32_t *trow = tile + (lane * R) * 33 + (kc & 31);
#pragma unroll
                for (int k = 0; k < R; k++) {
                    const int old_v = right_v[k];
                    int m;
                    if (AFFINE) {
                        const int old_d = right_d[k];
                        const int c0 = below_v + (below_d == GapX ? ext : open);
                        const int c1 = old_v + (old_d == GapY ? ext : open);
                        int d;
                        if (c0 < c1) {
                            m = c0;
                            d = GapX;
                        } else {
                            m = c1;
                            d = GapY;
                        }
                        const int c2 = dg + cst[k];
                        if (c2 < m) {
                            m = c2;
                            d = NoGap;
                        }
                        right_d[k] = d;
                        below_d = d;
                    } else {
                        m = __vimin3_s32(below_v + open, old_v + open, dg + cst[k]);
                    }
                    cst[k] = costrow[k][bcur]; // next column's substitution cost, off the critical path
                    right_v[k] = m;
                    trow[k * 33] = m;
                    dg = old_v;
                    below_v = m;
                }
                top_pk = (below_v << 2) | below_d;
                if (lane == 31 && has_consumer) ring_out[(out_base + kc) & (DP_RING - 1)] = top_pk;
            }
            // --- publish the top row's progress
            if (has_consumer && ((s & (DP_CHUNK - 1)) == DP_CHUNK - 1 || s == nsteps - 1)) {
                __threadfence_block();
                const int k31 = s - 31;
                if (lane == 31 && k31 >= 0) s_prod[warp] = out_base + (unsigned)min(k31 + 1, L2);
            }
            // --- write the staged tile out as row segments of up to 32 cells
            if ((s & 31) == 31 || s == nsteps - 1) {
                __syncwarp();
                TC *rowp = M + (size_t)(L1 - 1 - band * BAND) * cols + (L2 - 1); // row of rr = 0, column of kk = 0
                const int32_t *trd = tile;
                int rr = 0;
                for (int ln = 0; ln < 32 && rr < rows_here; ln++) {
                    const int kk = last_flush + 1 - ln + lane;
                    const bool ok = kk >= 0 && kk <= s - ln && kk < L2;
                    const int slot = kk & 31;
#pragma unroll
                    for (int k = 0; k < R; k++, rr++) {
                        if (ok && rr < rows_here) rowp[-kk] = (TC)trd[slot];
                        rowp -= cols;
                        trd += 33;
                    }
                }
                last_flush = s;
                __syncwarp();
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Linear-gap fast path (GapOpen == GapExtension, the reference's constants, Cost.h:13; all costs in 0..255).
//
// Same mapping as above (one CTA per pair, a warp per band of 32*R rows, lane = R rows, anti-diagonal wavefront, bands
// chained through shared-memory rings), but the step is cut to what the recurrence needs; the old kernel issued ~200
// warp instructions per step and was bound by instruction issue, not by the min-plus chain:
//   * the column residue travels WITH the wavefront: lane l-1 hands lane l one word {top value, residue code} by a
//     single __shfl_up; lane 0's inputs {cell of the row below the band, residue code} are fetched 16 steps at a time
//     (one lane each) and broadcast by __shfl;
//   * substitution costs come from a per-warp table T[code][lane] = the costs of the lane's R row residues against
//     residue `code`, packed one byte each: one conflict-free shared load per step for all R cells;
//   * a cell is min(min(below, right) + gap, diag + cost): VIMNMX + VIADDMNMX on the dependent chain;
//   * results go to a per-warp tile indexed by step (no per-lane address arithmetic) and leave as coalesced row
//     segments every 32 steps; 16-step blocks are fully unrolled, blocks in which every lane is inside the table run
//     without bounds checks.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FD_RING = 256; // ring entries per warp (power of two)
constexpr int FD_CHUNK = 16; // steps per unrolled block = ring fetch / publish granularity
constexpr int FD_TW = 32;    // tile width in steps = flush period
constexpr int FD_RS = FD_TW + 1;

template <typename TC, int R, bool CHECK>
__device__ __forceinline__ void fd_steps(int s0, int lane, int L2, int gap, const uint32_t *Tw, TC *tcol, int (&right)[R], int &diag0,
                                         int &out_pk, int in_pk, bool wr_ring, volatile int32_t *ring_out, unsigned out_base)
{
#pragma unroll
    for (int t = 0; t < FD_CHUNK; t++) {
        const int s = s0 + t;
        const int up = __shfl_up_sync(0xffffffffu, out_pk, 1);
        const int in0 = __shfl_sync(0xffffffffu, in_pk, t);
        const int cur = lane == 0 ? in0 : up;
        const int kc = s - lane; // my column counter: column L2-1-kc
        if (!CHECK || (unsigned)kc < (unsigned)L2) {
            const int code = cur & 127;
            int below = cur >> 7;
            const uint32_t cw = Tw[code * 32];
            int dg = diag0;
            diag0 = below;
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int c = (int)((cw >> (8 * k)) & 0xffu);
                const int m = __viaddmin_s32(min(below, right[k]), gap, dg + c);
                dg = right[k];
                right[k] = m;
                below = m;
                tcol[k * FD_RS + t] = (TC)m;
            }
            out_pk = (below << 7) | code;
            if (wr_ring) ring_out[(out_base + (unsigned)kc) & (FD_RING - 1)] = below;
        }
    }
}

template <typename TC, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1) pair_dp_linear_kernel(const __grid_constant__ DevProblem p, int warps, int nalpha)
{
    constexpr int R = 4, BAND = 32 * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *s_cost8 = smem_raw;                                                  // 90 x 90 costs, one byte each
    uint8_t *s_code = s_cost8 + 8112;                                             // residue -> dense code (96)
    uint8_t *s_alpha = s_code + 96;                                               // dense code -> residue (96)
    volatile int32_t *s_ring = reinterpret_cast<volatile int32_t *>(s_alpha + 96); // [warps][FD_RING]
    volatile unsigned *s_prod = reinterpret_cast<volatile unsigned *>(s_ring + warps * FD_RING);
    volatile unsigned *s_cons = s_prod + warps;
    uint32_t *s_T = const_cast<uint32_t *>(reinterpret_cast<volatile uint32_t *>(s_cons + warps)); // [warps][nalpha][32]
    TC *s_tile = reinterpret_cast<TC *>(s_T + (size_t)warps * nalpha * 32);                        // [warps][BAND][FD_RS]
    __shared__ int s_na;

    const int pair = blockIdx.x;
    const int sa = p.pa[pair], sb = p.pb[pair];
    const int L1 = p.len[sa], L2 = p.len[sb];
    const int cols = L2 + 1;
    TC *M = const_cast<TC *>(reinterpret_cast<const TC *>(p.table[pair]));
    const uint8_t *s1 = p.seq[sa];
    const uint8_t *s2 = p.seq[sb];
    const int gap = p.gap_open; // == p.gap_ext on this path

    for (int i = threadIdx.x; i < 90 * 90; i += blockDim.x) s_cost8[i] = (uint8_t)__ldg(p.cost + i);
    if (threadIdx.x < 96) s_code[threadIdx.x] = 0;
    if (threadIdx.x < warps) {
        s_prod[threadIdx.x] = 0;
        s_cons[threadIdx.x] = 0;
    }
    // borders, PairAlign.cpp:142-160
    for (int j = threadIdx.x; j <= L2; j += blockDim.x) M[(size_t)L1 * cols + j] = (TC)(j == L2 ? 0 : gap + (L2 - 1 - j) * gap);
    for (int i = threadIdx.x; i < L1; i += blockDim.x) M[(size_t)i * cols + L2] = (TC)(gap + (L1 - 1 - i) * gap);
    __syncthreads();
    for (int j = threadIdx.x; j < L2; j += blockDim.x) s_code[s2[j]] = 1; // residues that occur in the column sequence
    __syncthreads();
    if (threadIdx.x == 0) { // dense codes (at most nalpha: the host counted the distinct residues of all sequences)
        int na = 0;
        for (int b = 0; b < 90; b++)
            if (s_code[b]) {
                s_alpha[na] = (uint8_t)b;
                s_code[b] = (uint8_t)na;
                na++;
            }
        s_na = na;
    }
    __syncthreads();
    if (L2 == 0) return;
    const int na = s_na;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbands = (L1 + BAND - 1) / BAND;
    uint32_t *Tw = s_T + (size_t)warp * nalpha * 32 + lane;
    TC *tile = s_tile + (size_t)warp * BAND * FD_RS;
    volatile int32_t *ring_out = s_ring + warp * FD_RING;
    const int pwarp = (warp + warps - 1) % warps;
    volatile int32_t *ring_in = s_ring + pwarp * FD_RING;
    const int nsteps = L2 + 31;
    const int nchunks = (nsteps + FD_CHUNK - 1) / FD_CHUNK;

    for (int band = warp; band < nbands; band += warps) {
        // lane l owns rows r0, r0-1, ..., r0-(R-1); k = 0 is the lowest of them, lane 0 / k 0 the bottom row of the band
        const int r0 = L1 - 1 - (band * BAND + lane * R);
        int right[R];
        {
            int a[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int r = r0 - k;
                a[k] = r >= 0 ? (int)s1[r] : 0;
                right[k] = gap + (L1 - 1 - r) * gap; // M[r][L2]
            }
            for (int c = 0; c < na; c++) { // T[c][lane]: only this lane ever reads its column, no synchronisation needed
                const int b = s_alpha[c];
                uint32_t w = 0;
#pragma unroll
                for (int k = 0; k < R; k++) w |= (uint32_t)s_cost8[a[k] * 90 + b] << (8 * k);
                Tw[c * 32] = w;
            }
        }
        int diag0 = (r0 + 1 == L1) ? 0 : gap + (L1 - 2 - r0) * gap; // M[r0+1][L2]
        int out_pk = 0, in_pk = 0;
        const bool has_producer = band > 0, has_consumer = band + 1 < nbands;
        const unsigned in_base = has_producer ? (unsigned)((band - 1) / warps) * (unsigned)L2 : 0u;
        const unsigned out_base = (unsigned)(band / warps) * (unsigned)L2;
        const bool wr_ring = has_consumer && lane == 31;
        const int rows_here = min(BAND, L1 - band * BAND); // rows of this band that exist
        TC *tlane = tile + (size_t)(lane * R) * FD_RS;

        for (int ch = 0; ch < nchunks; ch++) {
            const int s0 = ch * FD_CHUNK;
            // ---- lane 0's inputs for the next FD_CHUNK steps: lane q < FD_CHUNK fetches column counter s0 + q
            if (s0 < L2) {
                const int kk = s0 + lane;
                unsigned need = 0;
                if (has_producer) {
                    need = in_base + (unsigned)min(s0 + FD_CHUNK, L2);
                    while ((int)(s_prod[pwarp] - need) < 0) { }
                    __threadfence_block();
                }
                if (lane < FD_CHUNK && kk < L2) {
                    const int code = s_code[s2[L2 - 1 - kk]];
                    const int v = has_producer ? ring_in[(in_base + (unsigned)kk) & (FD_RING - 1)] : gap + kk * gap; // M[L1][L2-1-kk]
                    in_pk = (v << 7) | code;
                }
                if (has_producer) {
                    __syncwarp();
                    if (lane == 0) s_cons[pwarp] = need;
                }
            }
            // ---- back-pressure: lane 31's next FD_CHUNK values must fit in the ring
            if (has_consumer) {
                const int k31 = s0 - 31;
                if (k31 + FD_CHUNK > 0 && k31 < L2) {
                    const unsigned top = out_base + (unsigned)min(k31 + FD_CHUNK, L2);
                    while ((int)(top - s_cons[warp]) > FD_RING) { }
                }
            }
            TC *tcol = tlane + (s0 & (FD_TW - 1));
            if (s0 >= 31 && s0 + FD_CHUNK - 1 < L2)
                fd_steps<TC, R, false>(s0, lane, L2, gap, Tw, tcol, right, diag0, out_pk, in_pk, wr_ring, ring_out, out_base);
            else
                fd_steps<TC, R, true>(s0, lane, L2, gap, Tw, tcol, right, diag0, out_pk, in_pk, wr_ring, ring_out, out_base);
            // ---- publish the top row's progress
            if (has_consumer) {
                __threadfence_block();
                const int k31 = s0 + FD_CHUNK - 1 - 31;
                if (lane == 31 && k31 >= 0) s_prod[warp] = out_base + (unsigned)min(k31 + 1, L2);
            }
            // ---- every FD_TW steps: write the staged tile out, one coalesced row segment per row.  Slot q of row
            //      (ln, k) holds step s = w0 + q, i.e. column counter s - ln of table row top - (ln * R + k).
            if ((s0 & (FD_TW - 1)) == FD_TW - FD_CHUNK || ch == nchunks - 1) {
                __syncwarp();
                const int w0 = s0 & ~(FD_TW - 1);
                const int s_end = s0 + FD_CHUNK - 1;
                const int s = w0 + lane;
                const int top = L1 - 1 - band * BAND; // table row of (ln 0, k 0)
                const TC *trd = tile + lane;
                TC *dst = M + (size_t)top * cols + (L2 - 1 - s);
                if (w0 >= 31 && w0 + FD_TW - 1 < L2 && s_end >= w0 + FD_TW - 1 && rows_here == BAND) {
#pragma unroll 4
                    for (int ln = 0; ln < 32; ln++) {
#pragma unroll
                        for (int k = 0; k < R; k++) {
                            dst[ln] = trd[0];
                            dst -= cols;
                            trd += FD_RS;
                        }
                    }
                } else {
                    int rr = 0;
                    for (int ln = 0; ln < 32 && rr < rows_here; ln++) {
                        const int kc = s - ln;
                        const bool ok = s <= s_end && kc >= 0 && kc < L2;
#pragma unroll
                        for (int k = 0; k < R; k++, rr++) {
                            if (ok && rr < rows_here) dst[ln] = trd[0];
                            dst -= cols;
                            trd += FD_RS;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
}

} // namespace

template <int DP_R, int DP_MAXW>
static int launch_pair_dp_cfg(pg_ctx *ctx, float *kernel_ms)
{
    int max_bands = 1;
    for (const PairGeom &g : ctx->pairs) max_bands = std::max(max_bands, (g.rows - 1 + 32 * DP_R - 1) / (32 * DP_R));
    const int warps = std::min(DP_MAXW, max_bands);
    const size_t smem = sizeof(int32_t) * 90 * 90 + (size_t)warps * DP_RING * 4 + (size_t)warps * 32 * DP_R * 33 * 4 + (size_t)warps * 8;
    const bool affine = ctx->dp.gap_open != ctx->dp.gap_ext;
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<ctx->npairs, warps * 32, smem, ctx->stream>>>(ctx->dp, warps);
        return cudaGetLastError();
    };
    cudaEvent_t e0, e1;
    PG_CUDA(ctx, cudaEventCreate(&e0));
    PG_CUDA(ctx, cudaEventCreate(&e1));
    PG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (ctx->dp.cell16) {
        if (affine)
            PG_CUDA(ctx, launch(pair_dp_kernel<uint16_t, true, DP_R, DP_MAXW>));
        else
            PG_CUDA(ctx, launch(pair_dp_kernel<uint16_t, false, DP_R, DP_MAXW>));
    } else {
        if (affine)
            PG_CUDA(ctx, launch(pair_dp_kernel<int32_t, true, DP_R, DP_MAXW>));
        else
            PG_CUDA(ctx, launch(pair_dp_kernel<int32_t, false, DP_R, DP_MAXW>));
    }
    PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    PG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    return PG_OK;
}

// Linear-gap fast path: see pair_dp_linear_kernel.
static int launch_pair_dp_linear(pg_ctx *ctx, float *kernel_ms)
{
    constexpr int MAXW = 8, R = 4;
    int max_bands = 1;
    for (const PairGeom &g : ctx->pairs) max_bands = std::max(max_bands, (g.rows - 1 + 32 * R - 1) / (32 * R));
    const int warps = std::min(MAXW, max_bands);
    const int nalpha = std::max(1, ctx->n_alpha);
    const size_t cell = ctx->dp.cell16 ? 2 : 4;
    const size_t smem = 8112 + 96 + 96 + (size_t)warps * FD_RING * 4 + (size_t)warps * 8 + (size_t)warps * nalpha * 32 * 4 +
                        (size_t)warps * 32 * R * FD_RS * cell + 16;
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<ctx->npairs, warps * 32, smem, ctx->stream>>>(ctx->dp, warps, nalpha);
        return cudaGetLastError();
    };
    cudaEvent_t e0, e1;
    PG_CUDA(ctx, cudaEventCreate(&e0));
    PG_CUDA(ctx, cudaEventCreate(&e1));
    PG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (ctx->dp.cell16)
        PG_CUDA(ctx, launch(pair_dp_linear_kernel<uint16_t, MAXW>));
    else
        PG_CUDA(ctx, launch(pair_dp_linear_kernel<int32_t, MAXW>));
    PG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    PG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    PG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    return PG_OK;
}

int pg_launch_pair_dp(pg_ctx *ctx, float *kernel_ms)
{
    // GapOpen == GapExtension (the reference's constants) and byte-sized costs: the lean linear-gap kernel.  Anything else
    // (general affine constants, costs beyond 255) runs the general kernel below.  PG_DP_KERNEL=general forces it.
    const char *force = getenv("PG_DP_KERNEL");
    if (ctx->dp.gap_open == ctx->dp.gap_ext && ctx->cost_u8 && ctx->dp.gap_open >= 0 && !(force && force[0] == 'g'))
        return launch_pair_dp_linear(ctx, kernel_ms);
    // measured on B200 (ms at S7 / S8): 8 rows/lane x 4 warps 0.209 / 0.421, 4 x 8 0.170 / 0.361, 2 x 16 0.190 / 0.430
    const char *e = getenv("PG_DP_CFG");
    const int cfg = e ? atoi(e) : 1;
    if (cfg == 0) return launch_pair_dp_cfg<8, 4>(ctx, kernel_ms);
    if (cfg == 2) return launch_pair_dp_cfg<2, 16>(ctx, kernel_ms);
    if (cfg == 3) return launch_pair_dp_cfg<1, 32>(ctx, kernel_ms);
    return launch_pair_dp_cfg<4, 8>(ctx, kernel_ms);
}
