// Host-side input producers of the hot path: the PAM250-as-cost table and the
// Altschul "rationale 2" pair weights.
//
// Reference: Cost::Cost (pastar/Cost.cpp:12-264) and weightAltschulsRationale2
// with primer / convert_path_to_cost / phylogeneticThreeNeighborJoin /
// compute_weights_from_tree (pastar/WeightedSP.cpp:424-519, 144-244, 109-142,
// 317-401, 403-420).  The reference runs this once per process on the host in
// single-precision float; g and h use the weights truncated to int
// (Node.cpp:226, HeuristicHPair.cpp:82), so every float operation here is done
// in the reference's order and type.  Build with -ffp-contract=off (the
// reference is built -O3 without FMA).  Not a GPU kernel on purpose: O(N^2 L^2)
// integers plus an O(N^5) float recursion whose value depends on operation order.
//
// The tree lives in index-addressed arrays instead of heap nodes, and the three
// alignment matrices are sized to the pair instead of the reference's fixed
// 1000x1000 scratch, which lifts its L <= 998 limit without changing any value.
#include <cstdint>
#include <cstring>
#include <string>
#include <algorithm>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

#include "pastar_gpu.h"

namespace {

// Dayhoff PAM250 log-odds scores, lower triangle, residue order as below.
// The reference's table is 17 - score (Cost.cpp:33-264), gaps vs 17 residues 12.
const char kResidues[] = "CSTPAGNDEQHRKMILVFYW";
const int8_t kPam250[20][20] = {
    {12},
    {0, 2},
    {-2, 1, 3},
    {-3, 1, 0, 6},
    {-2, 1, 1, 1, 2},
    {-3, 1, 0, -1, 1, 5},
    {-4, 1, 0, -1, 0, 0, 2},
    {-5, 0, 0, -1, 0, 1, 2, 4},
    {-5, 0, 0, -1, 0, 0, 1, 3, 4},
    {-5, -1, -1, 0, 0, -1, 1, 2, 2, 4},
    {-3, -1, -1, 0, -1, -2, 2, 1, 1, 3, 6},
    {-4, 0, -1, 0, -2, -3, 0, -1, -1, 1, 2, 6},
    {-5, 0, 0, -1, -1, -2, 1, 0, 0, 1, 0, 3, 5},
    {-5, -2, -1, -2, -1, -3, -2, -3, -2, -1, -2, 0, 0, 6},
    {-2, -1, 0, -2, -1, -3, -2, -2, -2, -2, -2, -2, -2, 2, 5},
    {-6, -3, -2, -3, -2, -4, -3, -4, -3, -2, -2, -3, -3, 4, 2, 6},
    {-2, -1, 0, -1, 0, -1, -2, -2, -2, -2, -2, -2, -2, 2, 4, 2, 4},
    {-4, -3, -3, -5, -4, -5, -4, -6, -5, -5, -2, -4, -5, 0, 1, 2, -1, 9},
    {0, -3, -3, -5, -3, -5, -2, -4, -4, -4, 0, -4, -4, -2, -1, -1, -2, 7, 10},
    {-8, -2, -5, -6, -6, -7, -4, -7, -7, -5, -3, 2, -3, -4, -5, -2, -6, 0, 0, 17},
};

struct CostTable {
    int32_t v[90 * 90];
    CostTable()
    {
        memset(v, 0, sizeof(v));
        for (int i = 0; i < 20; i++)
            for (int j = 0; j <= i; j++) {
                const int a = kResidues[i], b = kResidues[j];
                v[a * 90 + b] = v[b * 90 + a] = 17 - kPam250[i][j];
            }
        for (const char *c = "PAGNDEQHRKMILVFYW"; *c; ++c) v['-' * 90 + *c] = v[*c * 90 + '-'] = 12; // Cost.cpp:15-31
    }
    int operator()(unsigned char a, unsigned char b) const { return a < 90 && b < 90 ? v[a * 90 + b] : 0; }
};
const CostTable &cost_table()
{
    static const CostTable t;
    return t;
}

inline int min3(int a, int b, int c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
} // namespace

// identity-based distance x1000 of a pair from the number of identical columns on its alignment path
// (primer's tail, WeightedSP.cpp:222-228); en / em = the two sequence lengths
float pg_distance_from_matches(int en, int em, int match)
{
    const int scaled = (int)(0.5 + 1000.0 * (en - match + em - match) / (en + em));
    const float d = (float)scaled;
    return d <= 0 ? 1.0f : d; // WeightedSP.cpp:227-228
}
const int32_t *pg_cost_table_data();

namespace {

// Identity-based distance x1000 between two dash-prefixed strings (primer + convert_path_to_cost).
float pair_distance(const std::string &sa, const std::string &sb)
{
    const CostTable &cost = cost_table();
    const int kBig = 999999, kGap = 8, kEdgeGap = 0; // WeightedSP.hpp:13,18,22
    const int n = (int)sa.size(), m = (int)sb.size();
    const int W = m + 1;
    std::vector<int> dd((size_t)(n + 1) * W), hh((size_t)(n + 1) * W), vv((size_t)(n + 1) * W);
    auto at = [W](int i, int j) { return (size_t)i * W + j; };
    auto res = [](const std::string &s, int i) -> unsigned char { return i < (int)s.size() ? (unsigned char)s[i] : 0; };
    dd[0] = 0;
    hh[0] = vv[0] = kEdgeGap;
    for (int j = 1; j <= m; j++) {
        vv[at(0, j)] = dd[at(0, j)] = kBig;
        hh[at(0, j)] = hh[at(0, j - 1)] + cost('-', res(sb, j));
    }
    for (int i = 1; i <= n; i++) {
        hh[at(i, 0)] = dd[at(i, 0)] = kBig;
        vv[at(i, 0)] = vv[at(i - 1, 0)] + cost(res(sa, i), '-');
    }
    for (int i = 1; i < n; i++) {
        const int gi = i == n - 1 ? kEdgeGap : kGap;
        for (int j = 1; j < m; j++) {
            const int gj = j == m - 1 ? kEdgeGap : kGap;
            dd[at(i, j)] = min3(dd[at(i - 1, j - 1)], hh[at(i - 1, j - 1)], vv[at(i - 1, j - 1)]) + cost(res(sa, i), res(sb, j));
            hh[at(i, j)] = min3(dd[at(i, j - 1)] + gi, hh[at(i, j - 1)], vv[at(i, j - 1)] + gi) + cost('-', res(sb, j));
            vv[at(i, j)] = min3(dd[at(i - 1, j)] + gj, hh[at(i - 1, j)] + gj, vv[at(i - 1, j)]) + cost(res(sa, i), '-');
        }
    }
    // traceback from (n-1, m-1), counting identical columns (WeightedSP.cpp:109-142)
    const int en = n - 1, em = m - 1;
    enum { Diag, Vert, Horz };
    int dir = Diag, match = 0;
    for (int i = en, j = em; i || j;) {
        const int V = vv[at(i, j)] - (dir == Vert ? (j == em ? kEdgeGap : kGap) : 0);
        const int H = hh[at(i, j)] - (dir == Horz ? (i == en ? kEdgeGap : kGap) : 0);
        const int M = min3(V, H, dd[at(i, j)]);
        if (!j || M == V) {
            dir = Vert;
            --i;
        } else if (!i || M == H) {
            dir = Horz;
            --j;
        } else {
            dir = Diag;
            match += res(sa, i) == res(sb, j);
            --i;
            --j;
        }
    }
    return pg_distance_from_matches(en, em, match);
}

// Neighbour-joining tree in arrays.  kind: >= 0 leaf (sequence id), -1 internal, -2 root.
struct Tree {
    std::vector<int> kind, left, right, parent, brother;
    std::vector<float> len, w, W, v, V; // len = distance to parent ("weight" in the reference)
    int add(int k, int l, int r)
    {
        kind.push_back(k);
        left.push_back(l);
        right.push_back(r);
        parent.push_back(-1);
        brother.push_back(-1);
        len.push_back(0.0f);
        w.push_back(0.0f);
        W.push_back(0.0f);
        v.push_back(0.0f);
        V.push_back(0.0f);
        return (int)kind.size() - 1;
    }
};

struct NJ {
    Tree t;
    const std::vector<float> &D;
    int n;
    std::vector<int> live; // current forest roots ("tree" vector of the reference)
    NJ(const std::vector<float> &dist, int n_) : D(dist), n(n_) {}

    // sum of leaf-to-leaf distances between two subtrees, and the reference's hop count
    float cross(int a, int b, int &hops) const // compute_path_cost_rec, WeightedSP.cpp:248-266
    {
        if (t.kind[a] < 0) {
            ++hops;
            const float x = cross(t.left[a], b, hops);
            const float y = cross(t.right[a], b, hops);
            return x + y;
        }
        if (t.kind[b] < 0) {
            ++hops;
            const float x = cross(a, t.left[b], hops);
            const float y = cross(a, t.right[b], hops);
            return x + y;
        }
        return D[(size_t)t.kind[a] * n + t.kind[b]];
    }
    float mean_cross(int a, int b) const // compute_path_cost(_n), :270-288
    {
        int hops = 1;
        const float c = cross(a, b, hops);
        return (float)(c / hops);
    }
    float between(int i, int j) const { return mean_cross(live[i], live[j]); }
    float depth_sum(int a, float acc, int &count) const // compute_path_cost_to_leafs, :57-63
    {
        if (t.kind[a] >= 0) return acc + t.len[a];
        ++count;
        const float x = depth_sum(t.left[a], t.len[a] + acc, count);
        const float y = depth_sum(t.right[a], t.len[a] + acc, count);
        return x + y;
    }
    float branch(int i, int j) const // compute_curr_cost, :65-78
    {
        float di = 0.0f, dj = 0.0f;
        const int rem = (int)live.size();
        for (int k = 0; k < rem; k++)
            if (k != i && k != j) {
                di += between(i, k);
                dj += between(j, k);
            }
        di = di / (rem - 2);
        dj = dj / (rem - 2);
        const float dij = between(i, j);
        int count = 1;
        const float below = depth_sum(live[i], 0.0f, count); // call first, then divide by the updated count
        return (dij + di - dj) / 2 - below / count;
    }
    float criterion(int i, int j) const // compute_S, :290-311
    {
        const int rem = (int)live.size();
        float s1 = 0, s2 = 0;
        for (int k = 0; k < rem; k++)
            if (k != i && k != j) {
                const float a = between(i, k);
                const float b = between(j, k);
                s1 += a + b;
            }
        s1 = s1 / (2 * (rem - 2));
        for (int k = 0; k < rem - 1; k++)
            for (int l = k + 1; l < rem; l++)
                if (k != i && k != j && l != i && l != j) s2 += between(k, l);
        s2 = s2 / (rem - 2);
        return s1 + s2 + between(i, j) / 2;
    }
    void link(int parent, int l, int r)
    {
        t.brother[l] = r;
        t.brother[r] = l;
        t.parent[l] = t.parent[r] = parent;
    }
    void build() // phylogeneticThreeNeighborJoin + join_nodes, :317-401, 80-107
    {
        for (int i = 0; i < n; i++) live.push_back(t.add(i, -1, -1));
        while (live.size() > 2) {
            float best = (float)1.0E20;
            int bi = 0, bj = 0;
            const int rem = (int)live.size();
            for (int i = 0; i < rem - 1; i++)
                for (int j = i + 1; j < rem; j++) {
                    const float s = criterion(i, j);
                    if (s < best) {
                        best = s;
                        bi = i;
                        bj = j;
                    }
                }
            const int l = live[bi], r = live[bj];
            t.len[l] = branch(bi, bj);
            t.len[r] = branch(bj, bi);
            const int node = t.add(-1, l, r);
            link(node, l, r);
            live[bi] = node;
            live[bj] = live.back();
            live.pop_back();
        }
        const int l = live[0], r = live[1];
        const int root = t.add(-2, l, r);
        link(root, l, r);
        float len = mean_cross(l, r);
        int count = 1;
        const float dl = depth_sum(l, 0.0f, count);
        len -= dl / count;
        count = 1;
        const float dr = depth_sum(r, 0.0f, count);
        len -= dr / count;
        t.len[l] = len; // the right child of the root keeps its length (:397)
    }
    // compute_weights_from_tree, :403-420
    void spread(float product, float sum, int node, int brother, int from, std::vector<float> &out) const
    {
        if (t.kind[node] > -1) {
            out[(size_t)from * n + t.kind[node]] = sum * product;
        } else if (brother < 0) {
            const int l = t.left[node], r = t.right[node];
            spread(product * t.W[l], sum + t.len[r], r, -1, from, out);
            spread(product * t.W[r], sum + t.len[l], l, -1, from, out);
        } else {
            spread(product * t.V[node], sum + t.len[brother], brother, -1, from, out);
            if (t.kind[node] != -2) spread(product * t.W[brother], sum + t.len[node], t.parent[node], t.brother[node], from, out);
        }
    }
};

} // namespace

const int32_t *pg_cost_table_data() { return cost_table().v; }
int pg_weights_from_distances(int n, const std::vector<float> &dist, float *w_out);

extern "C" void pg_default_cost_table(int32_t out90x90[90 * 90]) { memcpy(out90x90, cost_table().v, sizeof(int32_t) * 8100); }

extern "C" int pg_host_weights(int n_seq, const char *const *seqs, const int *lens, float *w_out)
{
    if (n_seq < 2 || !seqs || !lens || !w_out) return PG_ERR_ARG;
    const int n = n_seq;
    std::vector<std::string> s(n);
    for (int i = 0; i < n; i++) {
        if (!seqs[i] || lens[i] < 1) return PG_ERR_ARG;
        for (int j = 0; j < lens[i]; j++) // 'Z' (90) and above index past pam250['Z']['Z'] in the reference (Cost.h:49): its
            if ((unsigned char)seqs[i][j] >= 90) return PG_ERR_ARG; // weights then vary with the environment; refused as in pg_ctx_create
        s[i] = "-" + std::string(seqs[i], seqs[i] + lens[i]); // WeightedSP.cpp:447
    }
    // The N(N-1)/2 pair distances (the reference's `primer`, WeightedSP.cpp:144-244) are independent of each other:
    // computed on as many host threads as there are pairs / cores.  Each pair runs the same float operations in the
    // same order as the serial loop, so the result is bit-identical.
    std::vector<float> dist((size_t)n * n, 0.0f);
    {
        std::vector<std::pair<int, int>> todo;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) todo.emplace_back(i, j);
        std::atomic<size_t> next(0);
        auto work = [&]() {
            for (size_t k = next.fetch_add(1); k < todo.size(); k = next.fetch_add(1)) {
                const int i = todo[k].first, j = todo[k].second;
                dist[(size_t)i * n + j] = dist[(size_t)j * n + i] = pair_distance(s[i], s[j]);
            }
        };
        unsigned nt = std::thread::hardware_concurrency();
        if (nt == 0) nt = 1;
        nt = (unsigned)std::min<size_t>(nt, todo.size());
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
    }

    return pg_weights_from_distances(n, dist, w_out);
}

// Neighbour-joining tree, partial weights, scaling (WeightedSP.cpp:317-420, 464-509) from the pair distances.  Shared by
// the host producer above and by pg_gpu_weights, whose device kernel produces the distances' match counts.
int pg_weights_from_distances(int n, const std::vector<float> &dist, float *w_out)
{
    NJ nj(dist, n);
    nj.build();
    Tree &t = nj.t;
    const int nodes = (int)t.kind.size(), root = nodes - 1;
    // partial weights, leaves first then internal nodes in creation order (:464-479)
    for (int k = 0; k < nodes; k++) {
        if (t.kind[k] > -1) {
            t.w[k] = 1.0f;
            t.W[k] = t.len[k];
        } else if (t.kind[k] > -2) {
            const int l = t.left[k], r = t.right[k];
            t.w[k] = t.w[l] * t.W[r] + t.W[l] * t.w[r];
            t.W[k] = t.len[k] * t.w[k] + t.W[l] * t.W[r];
        }
    }
    t.V[root] = 1;
    t.v[root] = 0;
    for (int k = root - 1; k >= 0; k--) { // (:485-492)
        const int p = t.parent[k], b = t.brother[k];
        t.v[k] = t.v[p] * t.W[b] + t.V[p] * t.w[b];
        t.V[k] = t.len[k] * t.v[k] + t.V[p] * t.W[b];
    }
    std::vector<float> raw((size_t)n * n, 0.0f);
    for (int k = 0; k < n; k++) nj.spread(1.0f, t.len[k], t.parent[k], t.brother[k], t.kind[k], raw);

    float smallest = 1.0E+30f; // scale so that the smallest weight is about 8 (:498-509)
    for (int j = 1; j < n; ++j)
        for (int i = 0; i < j; ++i)
            if (raw[(size_t)i * n + j] < smallest) smallest = raw[(size_t)i * n + j];
    smallest = (float)((double)smallest / 7.9);
    for (int i = 0; i < n * n; i++) w_out[i] = 0.0f;
    for (int i = 0; i < n - 1; ++i)
        for (int j = i + 1; j < n; ++j)
            w_out[(size_t)i * n + j] = w_out[(size_t)j * n + i] = (float)((double)(raw[(size_t)i * n + j] / smallest) + 0.5);
    return PG_OK;
}
