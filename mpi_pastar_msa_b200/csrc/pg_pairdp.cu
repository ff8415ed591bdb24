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
#include <cstddef>
#include <cstdlib>
#include <type_traits>

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
    const int cols = p.cols[pair]; // row pitch (>= L2 + 1)
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
// Same mapping as the general kernel (one CTA per pair, a warp per band of 32*R rows, anti-diagonal wavefront over the
// lanes, bands chained through shared-memory rings), rebuilt around what one warp can issue: the general kernel spends
// ~200 dependent warp instructions per single-column step.  Here
//   * a lane computes a register block of R x C = 4 x 4 cells per super-step (C columns of its R rows): 16 cells whose
//     dependency DAG is only R + C - 1 = 7 cells deep, so one lane-to-lane hand-over (C __shfl_up of the block's top
//     row, 24 cycles each, pipelined) is paid per FOUR columns and the cells in between overlap;
//   * a cell is min(min(below, right) + gap, diag + cost): PRMT (cost byte) + IADD + VIMNMX + VIADDMNMX;
//   * substitution costs come from a per-warp table T[code][lane] = the costs of the lane's R row residues against
//     column residue `code`, one byte each; the column codes sit in shared memory in sweep order (one aligned 32-bit
//     load gives a super-step's four codes); none of this is on the dependent chain;
//   * the table is swept on a 4-column grid aligned to memory: columns j > L2 are virtual +INF, the border column
//     j = L2 falls out of the recurrence itself, rows are pitched to a multiple of 8 cells, so every block leaves as
//     aligned 8 / 16-byte vectors: registers -> shared tile (STS.128) -> global (one vector per lane, 4 rows x 32
//     columns per warp instruction) every 8 super-steps;
//   * 4 super-steps are fully unrolled; chunks in which every lane is inside the table run without bounds checks.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FD_R = 4;       // rows per lane
constexpr int FD_C = 4;       // columns per super-step
constexpr int FD_SS = 4;      // super-steps per unrolled chunk (ring fetch / publish granularity: 16 columns)
constexpr int FD_WIN = 8;     // super-steps per flush window (32 columns)
constexpr int FD_RING = 256;  // ring entries (columns) per warp, power of two
constexpr int FD_CCPAD = 160; // padding of the column-code array on both sides (lanes ahead of / behind the table)
constexpr int FD_INF = 1 << 28;
// Ordering between a warp's ring accesses and its progress counters.  Both live in shared memory, and one warp's
// shared-memory accesses are performed in program order by the SM, so only the COMPILER must be kept from moving them:
// a MEMBAR.SC.CTA here (what __threadfence_block() emits) also waits for the flush's global stores - measured at about
// 1000 cycles, twice per 16 columns, 80 % of the first version's run time.
#define FD_ORDER() asm volatile("" ::: "memory")

template <typename TC>
struct FdGeom {
    static constexpr int UNIT = FD_R * FD_C * (int)sizeof(TC); // bytes one lane stages per super-step
    static constexpr int SG = 32 * UNIT + 16;                  // bytes per super-step slot of the tile (padded: bank spread)
    static constexpr int TILE = FD_WIN * SG;                   // bytes per warp
};

// predicated vector stores: a branch per store costs more than the store
__device__ __forceinline__ void fd_store_if(bool ok, void *ptr, const uint2 &v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v2.u32 [%1], {%2, %3};\n\t}" ::"r"((int)ok), "l"(ptr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fd_store_if(bool ok, void *ptr, const int4 &v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v4.u32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"((int)ok), "l"(ptr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// One chunk = FD_SS super-steps.  Every lane computes in every super-step: lanes ahead of the table (column group g < 0)
// or past it (g >= G) work on garbage that nothing valid ever reads - the wavefront hands a lane only blocks of the
// group it is about to compute - so the chunk has no bounds checks and no divergence.  RAMP (the first 32 super-steps of
// a band): a lane's running state is reset when it enters the table (g == 0).
//
// The sweep runs on N[i][j] = M[i][j] - gap * ((L1 - i) + (L2 - j)), the cost in excess of the all-gaps path.  Both gap
// moves then cost nothing and the diagonal costs c - 2 gap (tabulated as such), so a cell is ONE add and ONE three-input
// minimum,   N = min3(N_below, N_right, N_diag + c'),   and both borders are N = 0.  B200's integer pipes issue a warp
// instruction every other cycle, so the step is bound by its instruction count: costs are tabulated as ready-to-add
// 32-bit words (no byte extraction), M is restored when the block is staged (adj = gap * distance of its corner cell).
template <typename TC, bool RAMP>
__device__ __forceinline__ void fd_chunk(int u0, int lane, int G, int gap, const int4 *Tw, const uint8_t *cc_lane, unsigned char *tunit,
                                         const int *inp, int (&right)[FD_R], int &diag0, int (&top)[FD_C], int4 (&cwn)[FD_C], bool wr_ring,
                                         int *ring_out, unsigned out_pos, int adj0)
{
#pragma unroll
    for (int t = 0; t < FD_SS; t++) {
        int up[FD_C];
#pragma unroll
        for (int i = 0; i < FD_C; i++) up[i] = __shfl_up_sync(0xffffffffu, top[i], 1);
        const int4 in4 = *reinterpret_cast<const int4 *>(inp + FD_C * t); // lane 0's row below (uniform address: broadcast)
        const int g = u0 + t - lane;                                      // my column group: columns Lp-1-4g-i, i = 0..3
        int cw[FD_C][FD_R];
#pragma unroll
        for (int i = 0; i < FD_C; i++) {
            cw[i][0] = cwn[i].x;
            cw[i][1] = cwn[i].y;
            cw[i][2] = cwn[i].z;
            cw[i][3] = cwn[i].w;
        }
        {   // the next super-step's cost words travel while this block is computed
            const uint32_t codes = *reinterpret_cast<const uint32_t *>(cc_lane + FD_C * (t + 1));
#pragma unroll
            for (int i = 0; i < FD_C; i++) cwn[i] = Tw[((codes >> (8 * i)) & 0xffu) * 32];
        }
        if (RAMP && g == 0) {
#pragma unroll
            for (int k = 0; k < FD_R; k++) right[k] = FD_INF; // virtual column Lp
            diag0 = FD_INF;
        }
        int prev[FD_C];
        prev[0] = lane == 0 ? in4.x : up[0];
        prev[1] = lane == 0 ? in4.y : up[1];
        prev[2] = lane == 0 ? in4.z : up[2];
        prev[3] = lane == 0 ? in4.w : up[3];
        int prevR = diag0; // cell diagonally below-right of (row 0, column 0)
        diag0 = prev[FD_C - 1];
        int m[FD_R][FD_C];
#pragma unroll
        for (int k = 0; k < FD_R; k++) {
            int d = prevR, r = right[k];
            prevR = r;
#pragma unroll
            for (int i = 0; i < FD_C; i++) {
                const int v = __vimin3_s32(prev[i], r, d + cw[i][k]);
                d = prev[i];
                prev[i] = v;
                r = v;
                m[k][i] = v;
            }
            right[k] = r;
        }
#pragma unroll
        for (int i = 0; i < FD_C; i++) top[i] = prev[i];
        {   // back to M: cell (k, i) lies k + i steps further from the corner than the block's cell (0, 0)
            int e[FD_R + FD_C - 1];
            e[0] = adj0 + 4 * gap * t;
#pragma unroll
            for (int q = 1; q < FD_R + FD_C - 1; q++) e[q] = e[q - 1] + gap;
#pragma unroll
            for (int k = 0; k < FD_R; k++)
#pragma unroll
                for (int i = 0; i < FD_C; i++) m[k][i] += e[k + i];
        }
        // ---- stage the block: memory order is ascending column, i.e. i = 3, 2, 1, 0
        unsigned char *tu = tunit + t * FdGeom<TC>::SG;
        if constexpr (sizeof(TC) == 2) {
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < FD_R; k++) {
                w[2 * k] = __byte_perm((unsigned)m[k][3], (unsigned)m[k][2], 0x5410u);
                w[2 * k + 1] = __byte_perm((unsigned)m[k][1], (unsigned)m[k][0], 0x5410u);
            }
            reinterpret_cast<uint4 *>(tu)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4 *>(tu)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
#pragma unroll
            for (int k = 0; k < FD_R; k++) reinterpret_cast<int4 *>(tu)[k] = make_int4(m[k][3], m[k][2], m[k][1], m[k][0]);
        }
        if (wr_ring && (unsigned)g < (unsigned)G)
            *reinterpret_cast<int4 *>(ring_out + ((out_pos + (unsigned)(FD_C * g)) & (FD_RING - 1))) = make_int4(top[0], top[1], top[2], top[3]);
    }
}

template <typename TC, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1) pair_dp_linear_kernel(const __grid_constant__ DevProblem p, int warps, int nalpha, int cc_bytes)
{
    constexpr int R = FD_R, C = FD_C, BAND = 32 * R;
    using GEO = FdGeom<TC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *s_cost8 = smem_raw;                                   // 90 x 90 costs, one byte each (8112 reserved)
    uint8_t *s_code = s_cost8 + 8112;                              // residue -> dense code (96)
    uint8_t *s_alpha = s_code + 96;                                // dense code -> residue (96)
    int *s_ring = reinterpret_cast<int *>(s_alpha + 96);           // [warps][FD_RING]
    int *s_bord = s_ring + warps * FD_RING;                        // [warps][16]: band 0's row below (the border row)
    volatile unsigned *s_prod = reinterpret_cast<volatile unsigned *>(s_bord + warps * 16);
    volatile unsigned *s_cons = s_prod + warps;
    int4 *s_T = reinterpret_cast<int4 *>(const_cast<unsigned *>(s_cons + warps) + ((warps & 1) ? 2 : 0)); // [warps][nalpha][32] x 4 rows; 16-byte aligned
    uint8_t *s_cc = reinterpret_cast<uint8_t *>(s_T + (size_t)warps * nalpha * 32);                                          // column codes, sweep order
    unsigned char *s_tile = s_cc + cc_bytes;                                                                                  // [warps][GEO::TILE]
    __shared__ int s_na;

    const int pair = blockIdx.x;
    const int sa = p.pa[pair], sb = p.pb[pair];
    const int L1 = p.len[sa], L2 = p.len[sb];
    const int pitch = p.cols[pair];
    TC *M = const_cast<TC *>(reinterpret_cast<const TC *>(p.table[pair]));
    const uint8_t *s1 = p.seq[sa];
    const uint8_t *s2 = p.seq[sb];
    const int gap = p.gap_open;                 // == p.gap_ext on this path
    const int Lp = (L2 + 1 + C - 1) & ~(C - 1); // swept columns: 0 .. L2 (border column included), padded to the 4-column grid
    const int G = Lp / C;                       // column groups

    {   // cost table -> bytes (8100 = 2025 x 4 entries)
        const int4 *src = reinterpret_cast<const int4 *>(p.cost);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_cost8);
#pragma unroll 4
        for (int i = threadIdx.x; i < 2025; i += blockDim.x) {
            const int4 v = __ldg(src + i);
            dst[i] = (uint32_t)(v.x & 0xff) | ((uint32_t)(v.y & 0xff) << 8) | ((uint32_t)(v.z & 0xff) << 16) | ((uint32_t)(v.w & 0xff) << 24);
        }
    }
    for (int i = threadIdx.x; i < 96; i += blockDim.x) s_code[i] = 0;
    if (threadIdx.x < warps) {
        s_prod[threadIdx.x] = 0;
        s_cons[threadIdx.x] = 0;
    }
    // bottom border row M[L1][j] = gap * (L2 - j), PairAlign.cpp:142-160 (the border column comes out of the sweep)
    for (int j = threadIdx.x; j <= L2; j += blockDim.x) M[(size_t)L1 * pitch + j] = (TC)(gap * (L2 - j));
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
    // column codes in sweep order: entry FD_CCPAD + kc is column Lp-1-kc (0 for the border / virtual columns and the padding)
    for (int i = threadIdx.x; i < cc_bytes; i += blockDim.x) {
        const int j = Lp - 1 - (i - FD_CCPAD);
        s_cc[i] = (j >= 0 && j < L2) ? s_code[s2[j]] : (uint8_t)0;
    }
    __syncthreads();
    const int na = s_na;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbands = (L1 + BAND - 1) / BAND;
    int4 *Tw = s_T + (size_t)warp * nalpha * 32 + lane;
    unsigned char *tile = s_tile + (size_t)warp * GEO::TILE;
    int *ring_out = s_ring + warp * FD_RING;
    const int pwarp = (warp + warps - 1) % warps;
    const int *ring_in = s_ring + pwarp * FD_RING;
    int *bord = s_bord + warp * 16;
    const int nss = G + 31;                              // super-steps per band
    const int nchunks = (nss + FD_SS - 1) / FD_SS;
    const unsigned span = (unsigned)((Lp + 15) & ~15);   // ring positions per band (a chunk's 16 never wrap)

    for (int band = warp; band < nbands; band += warps) {
        // lane l owns rows r0, r0-1, ..., r0-(R-1); k = 0 is the lowest of them, lane 0 / k 0 the bottom row of the band
        const int r0 = L1 - 1 - (band * BAND + lane * R);
        int right[R], top[C];
        {
            int a[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int r = r0 - k;
                a[k] = r >= 0 ? (int)s1[r] : 0;
                right[k] = FD_INF;
            }
            for (int c = 0; c < na; c++) { // T[c][lane] = cost - 2 gap of the lane's 4 rows against residue code c.  Only this
                const int b = s_alpha[c];  // lane ever reads its column: no synchronisation needed
                Tw[c * 32] = make_int4((int)s_cost8[a[0] * 90 + b] - 2 * gap, (int)s_cost8[a[1] * 90 + b] - 2 * gap, (int)s_cost8[a[2] * 90 + b] - 2 * gap,
                                       (int)s_cost8[a[3] * 90 + b] - 2 * gap);
            }
        }
#pragma unroll
        for (int i = 0; i < C; i++) top[i] = 0;
        int diag0 = FD_INF;
        int4 cwn[C];
        {   // cost words of the first super-step
            const uint32_t codes = *reinterpret_cast<const uint32_t *>(s_cc + FD_CCPAD - C * lane);
#pragma unroll
            for (int i = 0; i < C; i++) cwn[i] = Tw[((codes >> (8 * i)) & 0xffu) * 32];
        }
        const bool has_producer = band > 0, has_consumer = band + 1 < nbands;
        const unsigned in_pos = has_producer ? (unsigned)((band - 1) / warps) * span : 0u;
        const unsigned out_pos = (unsigned)(band / warps) * span;
        const bool wr_ring = has_consumer && lane == 31;
        const int rows_here = min(BAND, L1 - band * BAND); // rows of this band that exist
        const int row_top = L1 - 1 - band * BAND;          // table row of (lane 0, k 0)

        for (int ch = 0; ch < nchunks; ch++) {
            const int u0 = ch * FD_SS;
            const int c0 = u0 * C; // first column counter of lane 0 in this chunk
            // ---- lane 0's row below for the chunk's 16 columns: the producer band's top row (ring) or the border row
            const int *inp = bord;
            if (has_producer) {
                if (c0 < Lp) {
                    const unsigned need = in_pos + (unsigned)min(c0 + FD_SS * C, Lp);
                    while ((int)(s_prod[pwarp] - need) < 0) { }
                    FD_ORDER();
                    if (lane == 0) s_cons[pwarp] = in_pos + (unsigned)c0; // everything before this chunk has been read
                }
                inp = ring_in + ((in_pos + (unsigned)c0) & (FD_RING - 1));
            } else {
                __syncwarp();
                if (lane < 16) {
                    const int j = Lp - 1 - (c0 + lane);
                    bord[lane] = (j >= 0 && j <= L2) ? 0 : FD_INF; // N = 0 on the border row
                }
                __syncwarp();
            }
            // ---- back-pressure: lane 31's next 16 columns must fit in the ring
            if (has_consumer) {
                const int g31 = u0 - 31; // lane 31's first group of this chunk
                if (g31 + FD_SS > 0 && g31 < G) {
                    const unsigned topw = out_pos + (unsigned)min((g31 + FD_SS) * C, Lp);
                    while ((int)(topw - s_cons[warp]) > FD_RING) { }
                }
            }
            unsigned char *tunit = tile + (u0 & (FD_WIN - 1)) * GEO::SG + lane * GEO::UNIT;
            const uint8_t *cc_lane = s_cc + FD_CCPAD + C * (u0 - lane);
            // gap * distance from the corner (L1, L2) of this lane's cell (k 0, i 0) at the chunk's first super-step:
            // row r0, column Lp - 1 - 4 (u0 - lane)
            const int adj0 = gap * ((L1 - r0) + (L2 - (Lp - 1 - C * (u0 - lane))));
            if (u0 < 32)
                fd_chunk<TC, true>(u0, lane, G, gap, Tw, cc_lane, tunit, inp, right, diag0, top, cwn, wr_ring, ring_out, out_pos, adj0);
            else
                fd_chunk<TC, false>(u0, lane, G, gap, Tw, cc_lane, tunit, inp, right, diag0, top, cwn, wr_ring, ring_out, out_pos, adj0);
            // ---- publish the top row's progress
            if (has_consumer) {
                FD_ORDER();
                const int g31 = u0 + FD_SS - 1 - 31; // lane 31's last finished group
                if (lane == 31 && g31 >= 0) s_prod[warp] = out_pos + (unsigned)min((g31 + 1) * C, Lp);
            }
            // ---- every FD_WIN super-steps: write the staged window out.  One 4-cell vector per lane: lanes 8 kq .. 8 kq + 7
            //      carry the window's 8 column groups of row ln * R + kq, i.e. 4 rows x 32 columns per warp instruction.
            if ((u0 & (FD_WIN - 1)) == FD_WIN - FD_SS || ch == nchunks - 1) {
                __syncwarp();
                const int w0 = u0 & ~(FD_WIN - 1);
                const int u_end = u0 + FD_SS - 1;
                const int sg = lane & 7, kq = lane >> 3;
                const int u = w0 + sg;
                const unsigned char *tsrc = tile + sg * GEO::SG + kq * (C * (int)sizeof(TC));
                // group of lane ln at super-step u: g = u - ln; its lowest column j0 = Lp - 4 g - 4; table row row_top - (ln R + kq)
                TC *dst = M + (ptrdiff_t)(row_top - kq) * pitch + (Lp - C * u - C);
                const ptrdiff_t step = C - (ptrdiff_t)R * pitch;
                // this lane's vectors are valid for ln in [lo, hi): its group u - ln inside the table, its row inside the band
                const int lnmax = (rows_here - kq + R - 1) / R;
                const int lo = max(0, u - G + 1);
                const int hi = u <= u_end ? min(min(32, u + 1), lnmax) : 0;
                const unsigned span_ln = (unsigned)max(hi - lo, 0);
                using V = typename std::conditional<sizeof(TC) == 2, uint2, int4>::type;
#pragma unroll 1
                for (int ln0 = 0; ln0 < 32; ln0 += 8) { // 8 vectors in flight: loads first, then predicated stores (no branches)
                    V buf[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) buf[q] = *reinterpret_cast<const V *>(tsrc + (ln0 + q) * GEO::UNIT);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int ln = ln0 + q;
                        fd_store_if((unsigned)(ln - lo) < span_ln, dst + ln * step, buf[q]);
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
    constexpr int MAXW = 8;
    int max_bands = 1, max_l2 = 1;
    for (const PairGeom &g : ctx->pairs) {
        max_bands = std::max(max_bands, (g.rows - 1 + 32 * FD_R - 1) / (32 * FD_R));
        max_l2 = std::max(max_l2, g.cols - 1);
    }
    int warps = std::min(MAXW, max_bands);
    const int nalpha = std::max(1, ctx->n_alpha);
    const int cc_bytes = (((max_l2 + 1 + FD_C - 1) & ~(FD_C - 1)) + 2 * FD_CCPAD + 15) & ~15;
    const size_t tile = ctx->dp.cell16 ? FdGeom<uint16_t>::TILE : FdGeom<int32_t>::TILE;
    auto smem_for = [&](int w) {
        return (size_t)8112 + 96 + 96 + (size_t)w * FD_RING * 4 + (size_t)w * 64 + (size_t)w * 8 + 8 + (size_t)w * nalpha * 32 * 16 + (size_t)cc_bytes +
               (size_t)w * tile + 16;
    };
    while (warps > 1 && smem_for(warps) > 220 * 1024) warps--; // long sequences / large alphabets: fewer bands in flight
    const size_t smem = smem_for(warps);
    if (smem > 227 * 1024) return -1; // does not fit: the caller runs the general kernel
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<ctx->npairs, warps * 32, smem, ctx->stream>>>(ctx->dp, warps, nalpha, cc_bytes);
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
    if (ctx->dp.gap_open == ctx->dp.gap_ext && ctx->cost_u8 && ctx->dp.gap_open >= 0 && !(force && force[0] == 'g')) {
        const int rc = launch_pair_dp_linear(ctx, kernel_ms);
        if (rc >= 0) return rc; // -1: the column codes of a very long sequence do not fit in shared memory
    }
    // measured on B200 (ms at S7 / S8): 8 rows/lane x 4 warps 0.209 / 0.421, 4 x 8 0.170 / 0.361, 2 x 16 0.190 / 0.430
    const char *e = getenv("PG_DP_CFG");
    const int cfg = e ? atoi(e) : 1;
    if (cfg == 0) return launch_pair_dp_cfg<8, 4>(ctx, kernel_ms);
    if (cfg == 2) return launch_pair_dp_cfg<2, 16>(ctx, kernel_ms);
    if (cfg == 3) return launch_pair_dp_cfg<1, 32>(ctx, kernel_ms);
    return launch_pair_dp_cfg<4, 8>(ctx, kernel_ms);
}
