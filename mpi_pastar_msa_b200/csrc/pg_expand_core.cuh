// Successor expansion core shared by the stand-alone getNeigh-batch kernel
// (pg_expand.cu) and the fused search kernel (pg_search.cu).
//
// Reference: Node<N>::getNeigh / pairCost / borderCheck / Node ctor
// (pastar/Node.cpp:205-248, 129-152, 69-77, 32-39), Coord<N>::neigh
// (pastar/Coord.cpp:92-106), HeuristicHPair::calculate_h
// (pastar/HeuristicHPair.cpp:73-86).
//
// The reference evaluates, for each of the 2^N-1 move masks, a sum over all
// P = N(N-1)/2 pairs of pairCost(mask)*w for g and a P-term table gather for
// h: S*P branchy terms and S*P dependent gathers per parent.  Both sums are
// quadratic pseudo-boolean functions of the mask bits whose coefficients
// depend on the parent only, so they are factorised here:
//
//   per pair a 4-entry LUT over (bit_x, bit_y):      4*P table gathers / parent
//       LUTg = {GapGap, gap(y), gap(x), cost(rx,ry)} * w   (Node.cpp:129-152)
//       LUTh = T[px+bx][py+by] * w                         (HeuristicHPair.cpp:82)
//   mask = (high << A) | low;  a group of 2^A lanes shares one parent, lane = low
//       B[low]    = sum over pairs inside low + cross pairs with by = 0
//       E_y[low]  = sum_x (LUT_xy[bx,1] - LUT_xy[bx,0])     one per high sequence y
//       HH[high]  = sum over pairs inside high              shared-memory table
//       value(low, high) = B[low] + HH[high] + sum_{y in high} E_y[low]
//
// so a successor costs two adds and two shared loads per sum instead of P
// terms.  All arithmetic is 32-bit signed integer, as in the reference.
#pragma once
#include "pg_internal.cuh"

template <int N>
struct ExpCfg {
    static constexpr int A = N <= 6 ? 3 : (N <= 9 ? 4 : 5); // mask bits enumerated across lanes
    static constexpr int LP = 1 << A;                       // lanes per parent
    static constexpr int HB = N - A;                        // mask bits enumerated in the per-lane loop
    static constexpr int IB = HB < 4 ? HB : 4;              // ... of which unrolled in registers
    static constexpr int UB = HB - IB;                      // ... and looped
    static constexpr int H = 1 << HB;
    static constexpr int P = N * (N - 1) / 2;
    static constexpr int S = (1 << N) - 1;
    static constexpr int NODE_WORDS = (((2 * N + 3) & ~3) + 12) / 4;
    static constexpr int POS_WORDS = ((2 * N + 3) & ~3) / 4;
    // ints of shared memory per parent group: LUTg, LUTh, HHg, HHh, pos
    static constexpr int GROUP_INTS = 8 * P + 2 * H + PG_MAX_SEQ;
};

// Per-CTA copy of the pair metadata (avoids divergent constant-bank indexing).
struct PairMeta {
    const void *table[PG_MAX_PAIRS];
    int w[PG_MAX_PAIRS];
    int cols[PG_MAX_PAIRS];
    uint8_t pa[PG_MAX_PAIRS], pb[PG_MAX_PAIRS];
};

__device__ __forceinline__ void pg_load_pair_meta(const DevProblem &p, PairMeta *m)
{
    for (int i = threadIdx.x; i < p.npairs; i += blockDim.x) {
        m->table[i] = p.table[i];
        m->w[i] = p.w[i];
        m->cols[i] = p.cols[i];
        m->pa[i] = p.pa[i];
        m->pb[i] = p.pb[i];
    }
}

// Expand one parent with a group of LP lanes.  `sub` is the lane's index in the
// group and its low mask bits, `gmask` the shfl mask of the group's lanes.
// s_grp points at GROUP_INTS ints of shared memory private to the group.
// The sink is called once per valid successor:
//   sink(mask, idx, posn, gnew, hnew)   idx = rank of mask among the valid masks (ascending)
template <int N, class Sink>
__device__ __forceinline__ void pg_expand_parent(const DevProblem &p, const PairMeta *meta, int *s_grp, const int (&pos)[N],
                                                 int g, int parenti, int sub, unsigned gmask, Sink &sink)
{
    using C = ExpCfg<N>;
    int *s_lutg = s_grp;
    int *s_luth = s_grp + 4 * C::P;
    int *s_hhg = s_grp + 8 * C::P;
    int *s_hhh = s_hhg + C::H;
    int *s_pos = s_hhh + C::H;

    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) s_pos[i] = pos[i];
    }
    int alive = 0; // sequences that can still advance (borderCheck, Node.cpp:69-77)
#pragma unroll
    for (int i = 0; i < N; i++) alive |= (pos[i] < p.len[i]) << i;
    __syncwarp(gmask);

    // ---- per-pair LUTs: 4*P entries, strided over the group's lanes
    for (int e = sub; e < 4 * C::P; e += C::LP) {
        const int pr = e >> 2, dx = e & 1, dy = (e >> 1) & 1;
        const int x = meta->pa[pr], y = meta->pb[pr];
        const int px = s_pos[x], py = s_pos[y];
        const int w = meta->w[pr];
        // moves that leave the lattice are never emitted; clamp so the gather stays in bounds
        const int ix = min(px + dx, p.len[x]), iy = min(py + dy, p.len[y]);
        const size_t cell = (size_t)ix * meta->cols[pr] + iy;
        const int t = p.cell16 ? (int)__ldg(reinterpret_cast<const uint16_t *>(meta->table[pr]) + cell)
                               : __ldg(reinterpret_cast<const int32_t *>(meta->table[pr]) + cell);
        int c;
        if (dx & dy)
            c = __ldg(p.cost + (int)__ldg(p.seq[x] + px) * 90 + (int)__ldg(p.seq[y] + py)); // Node.cpp:225
        else if (dx)
            c = ((parenti >> y) & 1) ? p.gap_open : p.gap_ext; // gap in y, Node.cpp:140,149-151
        else if (dy)
            c = ((parenti >> x) & 1) ? p.gap_open : p.gap_ext; // gap in x
        else
            c = p.gap_gap; // Node.cpp:142
        s_lutg[e] = c * w;
        s_luth[e] = t * w;
    }
    __syncwarp(gmask);

    // ---- HH[high]: pairs with both sequences in the high part
    for (int hi = sub; hi < C::H; hi += C::LP) {
        int sg = 0, sh = 0;
#pragma unroll
        for (int y = C::A; y < N; y++) {
#pragma unroll
            for (int z = y + 1; z < N; z++) {
                const int pr = y * N - y * (y + 1) / 2 + (z - y - 1);
                const int idx = pr * 4 + ((hi >> (y - C::A)) & 1) + 2 * ((hi >> (z - C::A)) & 1);
                sg += s_lutg[idx];
                sh += s_luth[idx];
            }
        }
        s_hhg[hi] = sg;
        s_hhh[hi] = sh;
    }

    // ---- B[low], E_y[low]
    int Bg = g, Bh = 0;
#pragma unroll
    for (int x = 0; x < C::A; x++) {
#pragma unroll
        for (int y = x + 1; y < C::A; y++) {
            const int pr = x * N - x * (x + 1) / 2 + (y - x - 1);
            const int idx = pr * 4 + ((sub >> x) & 1) + 2 * ((sub >> y) & 1);
            Bg += s_lutg[idx];
            Bh += s_luth[idx];
        }
    }
    int Eg[C::HB > 0 ? C::HB : 1], Eh[C::HB > 0 ? C::HB : 1];
#pragma unroll
    for (int y = C::A; y < N; y++) {
        int g0 = 0, g1 = 0, h0 = 0, h1 = 0;
#pragma unroll
        for (int x = 0; x < C::A; x++) {
            const int pr = x * N - x * (x + 1) / 2 + (y - x - 1);
            const int idx = pr * 4 + ((sub >> x) & 1);
            g0 += s_lutg[idx];
            g1 += s_lutg[idx + 2];
            h0 += s_luth[idx];
            h1 += s_luth[idx + 2];
        }
        Bg += g0;
        Bh += h0;
        Eg[y - C::A] = g1 - g0;
        Eh[y - C::A] = h1 - h0;
    }
    __syncwarp(gmask);

    // ---- enumerate the high bits
    int posn[N];
#pragma unroll
    for (int i = 0; i < C::A; i++) posn[i] = pos[i] + ((sub >> i) & 1);
    const int full = (1 << N) - 1;
    const bool interior = alive == full;

    for (int u = 0; u < (1 << C::UB); u++) {
        int vg[1 << C::IB], vh[1 << C::IB];
        vg[0] = Bg;
        vh[0] = Bh;
#pragma unroll
        for (int b = 0; b < C::UB; b++) {
            if ((u >> b) & 1) {
                vg[0] += Eg[C::IB + b];
                vh[0] += Eh[C::IB + b];
            }
            posn[C::A + C::IB + b] = pos[C::A + C::IB + b] + ((u >> b) & 1);
        }
#pragma unroll
        for (int b = 0; b < C::IB; b++) {
#pragma unroll
            for (int i = 0; i < (1 << b); i++) {
                vg[i + (1 << b)] = vg[i] + Eg[b];
                vh[i + (1 << b)] = vh[i] + Eh[b];
            }
        }
#pragma unroll
        for (int i = 0; i < (1 << C::IB); i++) {
            const int high = (u << C::IB) | i;
            const int mask = (high << C::A) | sub;
#pragma unroll
            for (int b = 0; b < C::IB; b++) posn[C::A + b] = pos[C::A + b] + ((i >> b) & 1);
            if (mask == 0) continue;
            if (!interior && (mask & ~alive)) continue;
            int idx = mask - 1;
            if (!interior) { // rank among the submasks of `alive`: compress the mask onto alive's bits
                int r = 0, o = 0;
#pragma unroll
                for (int b = 0; b < N; b++) {
                    if ((alive >> b) & 1) {
                        r |= ((mask >> b) & 1) << o;
                        o++;
                    }
                }
                idx = r - 1;
            }
            sink(mask, idx, posn, vg[i] + s_hhg[high], vh[i] + s_hhh[high]);
        }
    }
    __syncwarp(gmask);
}
