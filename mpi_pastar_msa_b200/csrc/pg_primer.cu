// The pair distances behind the Altschul weights, on the device.
//
// Reference: weightAltschulsRationale2 -> primer / convert_path_to_cost (pastar/WeightedSP.cpp:144-244, 109-142): for every
// pair of (dash-prefixed) sequences a forward three-matrix alignment (dd / hh / vv, GapCost 8, free end gaps), then a
// traceback from the last cell that counts identical columns; the distance is an integer function of that count.  The
// reference runs the N(N-1)/2 pairs serially on the host - half of HeuristicHPair::init.
//
// Here: one CTA per pair sweeps the anti-diagonals with one thread per row; the two previous diagonals of the three
// matrices live in shared memory (one barrier per diagonal).  The traceback never needs the matrix VALUES again, only,
// per cell, which move it would take for each of the three directions it can be entered with (the reference's
// comparison `M == V`, `M == H` with direction-dependent gap refunds): six bits, written once per cell as one byte.
// Thread 0 then walks the path through that byte map.  All arithmetic is integer, so the match count - and with it the
// distance - is exactly the reference's; the neighbour-joining tree and the float weight propagation stay on the host
// (pg_weights_from_distances: float-order exact), as SURVEY 2 scopes them.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pg_internal.cuh"

float pg_distance_from_matches(int en, int em, int match);                               // host/pg_host_weights.cpp
int pg_weights_from_distances(int n, const std::vector<float> &dist, float *w_out);      // host/pg_host_weights.cpp
const int32_t *pg_cost_table_data();                                                     // host/pg_host_weights.cpp

namespace {

constexpr int kBig = 999999, kGap = 8, kEdgeGap = 0; // WeightedSP.hpp:13,18,22
enum { Diag = 0, Vert = 1, Horz = 2 };

struct PrimerPair {
    const uint8_t *sa, *sb; // dash-prefixed residues, n and m of them, one trailing 0
    const int *pre_v;       // vv[i][0], i = 0..n (column 0: prefix sums of cost(sa[i], '-'))
    const int *pre_h;       // hh[0][j], j = 0..m
    uint8_t *dec;           // decision bytes, diagonal-major: cell (i, j) at dec[(i + j) * n + i] (coalesced along a diagonal)
    int n, m;
};

__device__ __forceinline__ int min3i(int a, int b, int c) { return min(a, min(b, c)); }

__global__ void __launch_bounds__(1024, 1) primer_kernel(const PrimerPair *pairs, const int32_t *cost, int maxn, int *match_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *s_cost = smem_raw;                              // 90 x 90 (the default table: every entry < 256)
    uint8_t *s_b = s_cost + 8112;                            // column residues, m + 1
    int *s_diag = reinterpret_cast<int *>(s_b + ((maxn + 1 + 15) & ~15)); // [3 diagonals][3 matrices][maxn]
    const PrimerPair P = pairs[blockIdx.x];
    const int n = P.n, m = P.m;
    for (int i = threadIdx.x; i < 8100; i += blockDim.x) s_cost[i] = (uint8_t)cost[i];
    for (int j = threadIdx.x; j <= m; j += blockDim.x) s_b[j] = j < m ? P.sb[j] : 0;
    __syncthreads();
    const int cvd_dash = '-' * 90;
    // cell (i, j) lies on diagonal d = i + j; the sweep covers rows 0 .. n-1 and columns 0 .. m-1 (primer's loops stop at
    // n-1, m-1: WeightedSP.cpp:206-220); row 0 and column 0 are the analytic borders
    const int last_d = (n - 1) + (m - 1);
    for (int d = 0; d <= last_d; d++) {
        int *cur = s_diag + (d % 3) * 3 * maxn;
        const int *p1 = s_diag + ((d + 2) % 3) * 3 * maxn; // diagonal d-1
        const int *p2 = s_diag + ((d + 1) % 3) * 3 * maxn; // diagonal d-2
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int j = d - i;
            if (j < 0 || j >= m) continue;
            int vdd, vhh, vvv;
            if (i == 0) { // WeightedSP.cpp:192-201
                vdd = j == 0 ? 0 : kBig;
                vvv = j == 0 ? kEdgeGap : kBig;
                vhh = P.pre_h[j];
            } else if (j == 0) {
                vdd = kBig;
                vhh = kBig;
                vvv = P.pre_v[i];
            } else {
                const int a = P.sa[i], b = s_b[j];
                const int gi = i == n - 1 ? kEdgeGap : kGap, gj = j == m - 1 ? kEdgeGap : kGap;
                // (i-1, j-1) on d-2 at row i-1; (i, j-1) on d-1 at row i; (i-1, j) on d-1 at row i-1
                vdd = min3i(p2[i - 1], p2[maxn + i - 1], p2[2 * maxn + i - 1]) + (int)s_cost[a * 90 + b];
                vhh = min3i(p1[i] + gi, p1[maxn + i], p1[2 * maxn + i] + gi) + (int)s_cost[cvd_dash + b];
                vvv = min3i(p1[i - 1] + gj, p1[maxn + i - 1] + gj, p1[2 * maxn + i - 1]) + (int)s_cost[a * 90 + '-'];
                // the move the traceback takes from here for each direction it can arrive with (WeightedSP.cpp:121-139)
                unsigned code = 0;
#pragma unroll
                for (int dir = 0; dir < 3; dir++) {
                    const int V = vvv - (dir == Vert ? (j == m - 1 ? kEdgeGap : kGap) : 0);
                    const int H = vhh - (dir == Horz ? (i == n - 1 ? kEdgeGap : kGap) : 0);
                    const int M = min3i(V, H, vdd);
                    const unsigned mv = M == V ? Vert : (M == H ? Horz : Diag);
                    code |= mv << (2 * dir);
                }
                P.dec[(size_t)d * n + i] = (uint8_t)code;
            }
            cur[i] = vdd;
            cur[maxn + i] = vhh;
            cur[2 * maxn + i] = vvv;
        }
        __syncthreads();
    }
    // ---- traceback from (n-1, m-1), counting identical columns (convert_path_to_cost, WeightedSP.cpp:109-142)
    if (threadIdx.x == 0) {
        int dir = Diag, match = 0;
        for (int i = n - 1, j = m - 1; i || j;) {
            int mv;
            if (!j)
                mv = Vert;
            else if (!i)
                mv = Horz; // row 0: vv is kBig, so M == V never holds there
            else
                mv = (P.dec[(size_t)(i + j) * n + i] >> (2 * dir)) & 3;
            if (mv == Vert) {
                --i;
            } else if (mv == Horz) {
                --j;
            } else {
                match += P.sa[i] == s_b[j];
                --i;
                --j;
            }
            dir = mv;
        }
        match_out[blockIdx.x] = match;
    }
}

} // namespace

// Replaces pg_host_weights' pair loop (the reference's primer, WeightedSP.cpp:144-244) by one kernel launch; same output,
// bit for bit.  Sequences longer than the shared-memory sweep supports return PG_ERR_UNSUPPORTED (callers use
// pg_host_weights then; the weight routine is a host-side input producer in the reference as well).
extern "C" int pg_gpu_weights(int n_seq, const char *const *seqs, const int *lens, int device, float *w_out, float *kernel_ms)
{
    if (n_seq < 2 || !seqs || !lens || !w_out) return PG_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PG_ERR_CUDA;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return PG_ERR_CUDA;
    const int n = n_seq;
    int maxn = 0;
    std::vector<std::string> s(n);
    for (int i = 0; i < n; i++) {
        if (!seqs[i] || lens[i] < 1) return PG_ERR_ARG;
        s[i] = "-" + std::string(seqs[i], seqs[i] + lens[i]); // WeightedSP.cpp:447
        for (unsigned char ch : s[i])
            if (ch >= 90) return PG_ERR_ARG;
        maxn = std::max(maxn, (int)s[i].size());
    }
    const size_t smem = 8112 + ((maxn + 1 + 15) & ~15) + (size_t)9 * maxn * sizeof(int);
    if (smem > 220 * 1024) return PG_ERR_UNSUPPORTED;
    const int32_t *cost = pg_cost_table_data();
    for (int i = 0; i < 8100; i++)
        if (cost[i] < 0 || cost[i] > 255) return PG_ERR_UNSUPPORTED;
    auto cst = [&](unsigned char a, unsigned char b) { return cost[a * 90 + b]; };

    // ---- host-side layout: residues, border prefix sums, per-pair descriptors
    std::vector<size_t> soff(n), voff(n), hoff(n);
    size_t sbytes = 0, ints = 0;
    for (int i = 0; i < n; i++) {
        soff[i] = sbytes;
        sbytes += (s[i].size() + 1 + 15) & ~size_t(15);
        voff[i] = ints;
        ints += s[i].size() + 1;
        hoff[i] = ints;
        ints += s[i].size() + 1;
    }
    std::vector<uint8_t> hseq(sbytes, 0);
    std::vector<int> hpre(ints, 0);
    for (int i = 0; i < n; i++) {
        memcpy(hseq.data() + soff[i], s[i].data(), s[i].size());
        const int len = (int)s[i].size();
        auto res = [&](int k) -> unsigned char { return k < len ? (unsigned char)s[i][k] : 0; };
        int *pv = hpre.data() + voff[i], *ph = hpre.data() + hoff[i];
        pv[0] = ph[0] = kEdgeGap;
        for (int k = 1; k <= len; k++) {
            pv[k] = pv[k - 1] + cst(res(k), '-'); // as the row sequence: vv[k][0]   (WeightedSP.cpp:198-201)
            ph[k] = ph[k - 1] + cst('-', res(k)); // as the column sequence: hh[0][k] (WeightedSP.cpp:192-196)
        }
    }
    const int npairs = n * (n - 1) / 2;
    size_t dec_bytes = 0;
    std::vector<size_t> doff(npairs);
    {
        int k = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++, k++) {
                doff[k] = dec_bytes;
                dec_bytes += ((s[i].size() + s[j].size()) * s[i].size() + 255) & ~size_t(255);
            }
    }
    // ---- ONE device allocation and ONE upload: [pairs | cost | prefix sums | residues | match counts | decision bytes]
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_pairs = 0, o_cost = up(npairs * sizeof(PrimerPair)), o_pre = o_cost + up(8100 * sizeof(int32_t)), o_seq = o_pre + up(ints * sizeof(int)),
                 o_match = o_seq + up(sbytes), o_dec = o_match + up(npairs * sizeof(int));
    const size_t staged = o_match, total = o_dec + dec_bytes;
    // the scratch buffer is kept per device between calls (grow-only): cudaMalloc + cudaFree of tens of MB cost more than
    // the kernel
    static std::mutex mu;
    static char *cache_ptr[64] = {nullptr};
    static size_t cache_bytes[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return PG_ERR_CUDA;
    if (cache_bytes[dev] < total) {
        if (cache_ptr[dev]) cudaFree(cache_ptr[dev]);
        cache_ptr[dev] = nullptr;
        cache_bytes[dev] = 0;
        if (cudaMalloc(&cache_ptr[dev], total) != cudaSuccess) return PG_ERR_CUDA;
        cache_bytes[dev] = total;
    }
    char *d_all = cache_ptr[dev];
    std::vector<char> stage(staged, 0);
    {
        PrimerPair *hp = reinterpret_cast<PrimerPair *>(stage.data() + o_pairs);
        int k = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++, k++) {
                hp[k].sa = reinterpret_cast<const uint8_t *>(d_all + o_seq + soff[i]);
                hp[k].sb = reinterpret_cast<const uint8_t *>(d_all + o_seq + soff[j]);
                hp[k].pre_v = reinterpret_cast<const int *>(d_all + o_pre) + voff[i];
                hp[k].pre_h = reinterpret_cast<const int *>(d_all + o_pre) + hoff[j];
                hp[k].dec = reinterpret_cast<uint8_t *>(d_all + o_dec + doff[k]);
                hp[k].n = (int)s[i].size();
                hp[k].m = (int)s[j].size();
            }
        memcpy(stage.data() + o_cost, cost, 8100 * sizeof(int32_t));
        memcpy(stage.data() + o_pre, hpre.data(), ints * sizeof(int));
        memcpy(stage.data() + o_seq, hseq.data(), sbytes);
    }
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) {
        if (e == cudaSuccess) e = x;
        return e == cudaSuccess;
    };
    std::vector<int> match(npairs, 0);
    float ms = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static int attr_smem[64] = {0}; // per device: the largest dynamic shared memory size opted in to so far
    ok(cudaMemcpy(d_all, stage.data(), staged, cudaMemcpyHostToDevice));
    if (e == cudaSuccess && attr_smem[dev] < (int)smem) {
        ok(cudaFuncSetAttribute(primer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[dev] = (int)smem;
    }
    if (kernel_ms) {
        ok(cudaEventCreate(&e0));
        ok(cudaEventCreate(&e1));
        ok(cudaEventRecord(e0));
    }
    if (e == cudaSuccess) {
        const int threads = std::min(1024, (maxn + 31) & ~31);
        primer_kernel<<<npairs, threads, smem>>>(reinterpret_cast<const PrimerPair *>(d_all + o_pairs), reinterpret_cast<const int32_t *>(d_all + o_cost), maxn,
                                                reinterpret_cast<int *>(d_all + o_match));
        ok(cudaGetLastError());
    }
    if (kernel_ms) ok(cudaEventRecord(e1));
    ok(cudaMemcpy(match.data(), d_all + o_match, npairs * sizeof(int), cudaMemcpyDeviceToHost));
    if (kernel_ms && e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) return PG_ERR_CUDA;
    if (kernel_ms) *kernel_ms = ms;
    std::vector<float> dist((size_t)n * n, 0.0f);
    {
        int k = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++, k++)
                dist[(size_t)i * n + j] = dist[(size_t)j * n + i] = pg_distance_from_matches((int)s[i].size() - 1, (int)s[j].size() - 1, match[k]);
    }
    return pg_weights_from_distances(n, dist, w_out);
}
