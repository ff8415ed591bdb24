// Kernel 2 — batched successor expansion (stand-alone getNeigh batch), plus the
// small calculate_h and get_id batch kernels.
//
// Replaces Node<N>::getNeigh (reference pastar/Node.cpp:205-248) for K parents
// per launch: reads K Node<N>-layout records, writes for each parent its valid
// successors {pos, f, g, parenti, owner} in ascending move-mask order at
// out[k * (2^N-1) ...] and the count at counts[k].
//
// Bound: HBM write bandwidth.  Algorithmic bytes per expansion:
//   B_parent + S * B_succ = sizeof(Node<N>) + N + 16*P  +  (2^N-1) * (sizeof(Node<N>) + 4)
//   = 4435 B at N=7, 8644 B at N=8.  The 4*P table gathers hit L2.
// Layout: a group of 2^A lanes (A = 3/4/5) expands one parent; at a fixed high
// mask the group's lanes write 2^A consecutive records, so every store
// instruction of a warp covers 32/2^A contiguous runs of 2^A * stride bytes.
#include "pg_expand_core.cuh"

namespace {

// Owner of a successor without rebuilding its Morton word.  For FZORDER / PZORDER the reference keeps
// nb = floor(log2(size)) + 2 Morton bits starting at bit `shift`; Morton bit (shift + m) is bit
// cb[m] of coordinate ci[m].  A successor differs from its parent by +1 on the coordinates in its mask, so per
// parent we compute W0 (the word of the parent itself) and, per coordinate i, D[i] = the bits that flip when
// coordinate i advances; the successor's word is W0 ^ XOR_{i in mask} D[i]  (each coordinate owns distinct bits).
struct OwnerPlan {
    int type, shift, size, nb, pow2;
    int ci[8], cb[8];
};

template <int N>
struct OwnerState {
    unsigned w0;   // z-order: Morton word of the parent; sum hashes: the parent's sum
    unsigned d[N]; // z-order: flip masks
};

template <int N>
__device__ __forceinline__ void owner_prepare(const OwnerPlan &op, const int (&pos)[N], OwnerState<N> &os)
{
    os.w0 = 0;
#pragma unroll
    for (int i = 0; i < N; i++) os.d[i] = 0;
    if (op.size <= 1) return;
    if (op.type == PG_HASH_FSUM || op.type == PG_HASH_PSUM) {
        const int nd = op.type == PG_HASH_PSUM ? 2 : N;
#pragma unroll
        for (int i = 0; i < N; i++)
            if (i < nd) os.w0 += (unsigned)pos[i];
        return;
    }
    for (int m = 0; m < op.nb; m++) {
        const int c = op.ci[m], b = op.cb[m];
        if (b >= 16) continue; // Coord is uint16: higher bits are always 0
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i == c) {
                const unsigned b0 = ((unsigned)pos[i] >> b) & 1u, b1 = ((unsigned)(pos[i] + 1) >> b) & 1u;
                os.w0 |= b0 << m;
                os.d[i] |= (b0 ^ b1) << m;
            }
        }
    }
}

template <int N>
__device__ __forceinline__ uint32_t owner_of_mask(const OwnerPlan &op, const OwnerState<N> &os, int mask)
{
    if (op.size <= 1) return 0;
    unsigned w;
    if (op.type == PG_HASH_FSUM) {
        w = (os.w0 + (unsigned)__popc(mask)) >> op.shift; // CoordHash.cpp:27-44
    } else if (op.type == PG_HASH_PSUM) {
        w = (os.w0 + (unsigned)__popc(mask & 3)) >> op.shift; // CoordHash.cpp:47-61
    } else {
        w = os.w0;
#pragma unroll
        for (int i = 0; i < N; i++)
            if ((mask >> i) & 1) w ^= os.d[i];
    }
    return op.pow2 ? (w & (unsigned)(op.size - 1)) : (w % (unsigned)op.size);
}

template <int N>
struct RecordSink {
    using C = ExpCfg<N>;
    uint32_t *out;   // this parent's first record
    const OwnerPlan &op;
    const OwnerState<N> &os;
    __device__ __forceinline__ void operator()(int mask, int idx, const int (&posn)[N], int gnew, int hnew)
    {
        constexpr int SW = C::POS_WORDS + 4; // words per successor record
        uint32_t w[SW];
#pragma unroll
        for (int j = 0; j < C::POS_WORDS; j++) {
            uint32_t lo = 2 * j < N ? (uint32_t)posn[2 * j] : 0u;
            uint32_t hi = 2 * j + 1 < N ? (uint32_t)posn[2 * j + 1] : 0u;
            w[j] = lo | (hi << 16);
        }
        w[C::POS_WORDS + 0] = (uint32_t)(gnew + hnew); // m_f, Node.cpp:38
        w[C::POS_WORDS + 1] = (uint32_t)gnew;
        w[C::POS_WORDS + 2] = (uint32_t)mask;          // parenti, Node.cpp:244
        w[C::POS_WORDS + 3] = owner_of_mask<N>(op, os, mask);
        uint32_t *dst = out + (size_t)idx * SW;
        if constexpr (SW % 4 == 0) {
#pragma unroll
            for (int j = 0; j < SW; j += 4) {
                // streaming 128-bit stores: records are written once and not re-read by this kernel
                asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "r"(w[j]), "r"(w[j + 1]),
                             "r"(w[j + 2]), "r"(w[j + 3])
                             : "memory");
            }
        } else {
#pragma unroll
            for (int j = 0; j < SW; j++) __stcs(dst + j, w[j]);
        }
    }
};

template <int N>
__global__ void __launch_bounds__(256, 3) expand_batch_kernel(const __grid_constant__ DevProblem p, const uint32_t *__restrict__ parents,
                                                           long long k, uint32_t *__restrict__ out, int32_t *__restrict__ counts,
                                                           const __grid_constant__ OwnerPlan op)
{
    using C = ExpCfg<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairMeta *meta = reinterpret_cast<PairMeta *>(smem_raw);
    int *s_groups = reinterpret_cast<int *>(smem_raw + ((sizeof(PairMeta) + 15) & ~size_t(15)));
    pg_load_pair_meta(p, meta);
    __syncthreads();

    constexpr int GROUPS = 256 / C::LP;
    const int grp = threadIdx.x / C::LP, sub = threadIdx.x % C::LP;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = C::LP == 32 ? 0xffffffffu : (((1u << C::LP) - 1u) << (lane & ~(C::LP - 1)));
    int *s_grp = s_groups + grp * C::GROUP_INTS;
    const bool aligned16 = true;
    (void)aligned16;

    for (long long base = (long long)blockIdx.x * GROUPS; base < k; base += (long long)gridDim.x * GROUPS) {
        const long long pi = base + grp;
        if (pi >= k) continue; // whole group leaves together
        // ---- load the parent record: lane j of the group reads word j, then broadcast
        uint32_t myw = sub < C::NODE_WORDS ? __ldg(parents + pi * C::NODE_WORDS + sub) : 0u;
        int pos[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t wv = __shfl_sync(gmask, myw, i / 2, C::LP);
            pos[i] = (i & 1) ? (int)(wv >> 16) : (int)(wv & 0xffffu);
        }
        const int g = (int)__shfl_sync(gmask, myw, C::POS_WORDS + 1, C::LP);
        const int parenti = (int)__shfl_sync(gmask, myw, C::POS_WORDS + 2, C::LP);

        OwnerState<N> os;
        owner_prepare<N>(op, pos, os);
        RecordSink<N> sink{out + (size_t)pi * C::S * (C::POS_WORDS + 4), op, os};
        pg_expand_parent<N>(p, meta, s_grp, pos, g, parenti, sub, gmask, sink);
        if (sub == 0) {
            int alive = 0;
#pragma unroll
            for (int i = 0; i < N; i++) alive += pos[i] < p.len[i];
            counts[pi] = (1 << alive) - 1;
        }
    }
}

// HeuristicHPair::calculate_h<N> (pastar/HeuristicHPair.cpp:73-86) for a batch of coords.
__global__ void calc_h_kernel(const __grid_constant__ DevProblem p, const uint16_t *__restrict__ coords, long long n,
                              int32_t *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint16_t *c = coords + i * p.n;
        int h = 0;
        for (int pr = 0; pr < p.npairs; pr++) h += pg_table_cell(p, pr, c[p.pa[pr]], c[p.pb[pr]]) * p.w[pr];
        out[i] = h;
    }
}

// Coord<N>::get_id (pastar/CoordHash.cpp:190-245) for a batch of coords.
__global__ void owner_kernel(const uint16_t *__restrict__ coords, long long n, int nseq, int hash_type, int shift, int size,
                             int log2size, uint32_t *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = pg_owner_of(coords + i * nseq, nseq, hash_type, shift, size, log2size);
}

int ilog2(int v)
{
    int l = 0;
    while ((1 << (l + 1)) <= v) l++;
    return l;
}

OwnerPlan make_owner_plan(const pg_ctx *ctx, int size)
{
    OwnerPlan op;
    op.type = ctx->dp.hash_type;
    op.shift = ctx->dp.hash_shift;
    op.size = size < 1 ? 1 : size;
    op.pow2 = (op.size & (op.size - 1)) == 0;
    op.nb = ilog2(op.size) + 2;
    if (op.nb > 8) op.nb = 8;
    const int nd = op.type == PG_HASH_PZORDER ? 2 : ctx->n;
    for (int m = 0; m < 8; m++) {
        const int q = op.shift + m;
        op.ci[m] = q % nd;
        op.cb[m] = q / nd;
    }
    return op;
}

template <int N>
int launch_expand_n(pg_ctx *ctx, const void *d_parents, int64_t k, int vec_size, void *d_out, int32_t *d_counts, cudaStream_t st)
{
    using C = ExpCfg<N>;
    constexpr int GROUPS = 256 / C::LP;
    const size_t smem = ((sizeof(PairMeta) + 15) & ~size_t(15)) + sizeof(int) * (size_t)GROUPS * C::GROUP_INTS;
    // the dynamic shared memory opt-in is per DEVICE state: cached in the context (one context per device), never in a
    // process-wide static (pg_multi_search drives several devices from one process)
    int &occ = ctx->occ_expand_batch;
    if (!occ) {
        PG_CUDA(ctx, cudaFuncSetAttribute(expand_batch_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, expand_batch_kernel<N>, 256, smem));
        if (occ < 1) occ = 1;
    }
    long long want = (k + GROUPS - 1) / GROUPS;
    long long grid = std::min<long long>(want, (long long)ctx->sm_count * occ); // persistent: a multiple of the SM count
    if (grid < 1) grid = 1;
    OwnerPlan op = make_owner_plan(ctx, vec_size);
    expand_batch_kernel<N><<<(unsigned)grid, 256, smem, st>>>(ctx->dp, reinterpret_cast<const uint32_t *>(d_parents), (long long)k,
                                                             reinterpret_cast<uint32_t *>(d_out), d_counts, op);
    PG_CUDA(ctx, cudaGetLastError());
    return PG_OK;
}

} // namespace

int pg_launch_expand(pg_ctx *ctx, const void *d_parents, int64_t k, int vec_size, void *d_out, int32_t *d_counts, cudaStream_t st)
{
    if (k <= 0) return PG_OK;
    switch (ctx->n) {
#define CASE(X) \
    case X:     \
        return launch_expand_n<X>(ctx, d_parents, k, vec_size, d_out, d_counts, st);
        CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(14) CASE(16)
        CASE(11) CASE(12) CASE(13) CASE(15) // pg_allow_extended_n
#undef CASE
    default:
        return pg_fail(ctx, PG_ERR_ARG, "unsupported number of sequences (reference supports 3-10, 14, 16: max_seq_helper.h:9-19)");
    }
}

int pg_launch_calc_h(pg_ctx *ctx, const uint16_t *d_coords, int64_t n, int32_t *d_out, cudaStream_t st)
{
    if (n <= 0) return PG_OK;
    long long grid = std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
    calc_h_kernel<<<(unsigned)grid, 256, 0, st>>>(ctx->dp, d_coords, (long long)n, d_out);
    PG_CUDA(ctx, cudaGetLastError());
    return PG_OK;
}

int pg_launch_owner(pg_ctx *ctx, const uint16_t *d_coords, int64_t n, int size, uint32_t *d_out, cudaStream_t st)
{
    if (n <= 0) return PG_OK;
    long long grid = std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
    owner_kernel<<<(unsigned)grid, 256, 0, st>>>(d_coords, (long long)n, ctx->n, ctx->dp.hash_type, ctx->dp.hash_shift, size,
                                                 ilog2(size < 1 ? 1 : size), d_out);
    PG_CUDA(ctx, cudaGetLastError());
    return PG_OK;
}
