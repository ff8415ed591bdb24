// TEST INFRASTRUCTURE — not part of the product.
//
// Driver around the UNMODIFIED reference sources (compiled in place from
// /root/reference/pastar by oracle/Makefile into oracle/_ref/).  It links the
// reference's own Cost, PairAlign, HeuristicHPair, WeightedSP, Coord,
// CoordHash, Sequences and Node objects and only adds what cannot be built in
// this image (Boost.MultiIndex / MPI / LZ4 are absent):
//
//   * a pos-unique + f-ordered open list with the semantics of
//     PriorityList.h:84-122 (dequeue / conditional_enqueue / highest priority),
//   * a serial A* loop following AStar.cpp:53-104,
//   * a T-thread, hash-partitioned PA-Star loop following
//     PAStar.cpp:219-237 (enqueue), :319-401 (worker_inner) and
//     :410-547 (process_final_node / check_stop), one process (R = 1).
//
// Everything numeric (tables, weights, g/h/f, owner ids) comes from the
// reference's own functions.  The Sequences singleton cannot be reset
// (Sequences.cpp:30-36 leaves `seqs` populated), so this is a one-problem-per-
// process executable; tests and bench.py run it as a subprocess.
//
// Output files are little-endian binary, documented at each writer below and
// parsed by oracle/refio.py.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "include/Coord.h"
#include "include/Cost.h"
#include "include/HeuristicHPair.h"
#include "include/Node.h"
#include "include/PairAlign.h"
#include "include/Sequences.h"
#include "include/max_seq_helper.h"
#include "include/read_fasta.h"
#include "include/backtrace.h"
#include <list>
#include <unistd.h>

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void put_i32(FILE *f, int32_t v) { fwrite(&v, 4, 1, f); }
void put_i64(FILE *f, int64_t v) { fwrite(&v, 8, 1, f); }
void put_f32(FILE *f, float v) { fwrite(&v, 4, 1, f); }

hashType parse_hash(const std::string &s)
{
    if (s == "FZORDER") return HashFZorder;
    if (s == "PZORDER") return HashPZorder;
    if (s == "FSUM") return HashFSum;
    if (s == "PSUM") return HashPSum;
    std::cerr << "bad hash type " << s << std::endl;
    exit(2);
}

// The reference's PairAlign objects are private to HeuristicHPair; rebuild
// them with the same constructor, in the same (i<j) order as
// HeuristicHPair.cpp:54-61, to read whole tables.
std::vector<PairAlign *> build_pairs()
{
    std::vector<PairAlign *> v;
    Sequences *seq = Sequences::getInstance();
    int n = Sequences::get_seq_num();
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++)
            v.push_back(new PairAlign(Pair(i, j), seq->get_seq(i), seq->get_seq(j)));
    return v;
}

// ---------------------------------------------------------------------------
// dump: N, lens, cost table, weights, all pairwise tables
//   magic "PGD1", i32 N, i32 len[N], i32 cost[90*90], f32 w[N*N],
//   for each pair (i<j): i32 rows, i32 cols, i32 cells[rows*cols]
// ---------------------------------------------------------------------------
int cmd_dump(const char *out)
{
    FILE *f = fopen(out, "wb");
    if (!f) return 3;
    Sequences *seq = Sequences::getInstance();
    int n = Sequences::get_seq_num();
    fwrite("PGD1", 4, 1, f);
    put_i32(f, n);
    for (int i = 0; i < n; i++) put_i32(f, (int)seq->get_seq(i).length());
    for (int a = 0; a < 90; a++)
        for (int b = 0; b < 90; b++) put_i32(f, Cost::cost((char)a, (char)b));
    HeuristicHPair *hp = HeuristicHPair::getInstance();
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) put_f32(f, hp->weightMatrix[i][j]);
    std::vector<PairAlign *> pairs = build_pairs();
    for (size_t p = 0; p < pairs.size(); p++) {
        int a = pairs[p]->getPair().first, b = pairs[p]->getPair().second;
        int rows = (int)seq->get_seq(a).length() + 1, cols = (int)seq->get_seq(b).length() + 1;
        put_i32(f, rows);
        put_i32(f, cols);
        std::vector<int32_t> row(cols);
        for (int i = 0; i < rows; i++) {
            for (int j = 0; j < cols; j++) row[j] = pairs[p]->getScore(i, j);
            fwrite(row.data(), 4, cols, f);
        }
    }
    fclose(f);
    return 0;
}

// ---------------------------------------------------------------------------
// neigh: run Node<N>::getNeigh on parents read from a file.
//   in : i32 K, then K x { u16 pos[N], i32 g, i32 parenti }   (packed)
//   out: "PGN1", i32 N, i32 K, then per parent: i32 f_parent, i32 count,
//        count x { u16 pos[N], i32 f, i32 g, i32 parenti, i32 owner }
//   Successor order is the reference's: bucket 0..vec_size-1, ascending mask
//   inside a bucket (Node.cpp:234-246).
// ---------------------------------------------------------------------------
template <int N>
int neigh_run(const char *in, const char *out, int vec_size)
{
    FILE *fi = fopen(in, "rb");
    FILE *fo = fopen(out, "wb");
    if (!fi || !fo) return 3;
    int32_t K = 0;
    if (fread(&K, 4, 1, fi) != 1) return 3;
    fwrite("PGN1", 4, 1, fo);
    put_i32(fo, N);
    put_i32(fo, K);
    std::vector<std::vector<Node<N> > > buckets(vec_size);
    for (int k = 0; k < K; k++) {
        uint16_t pos[N];
        int32_t g, parenti;
        if (fread(pos, 2, N, fi) != (size_t)N || fread(&g, 4, 1, fi) != 1 || fread(&parenti, 4, 1, fi) != 1) return 3;
        Coord<N> c;
        for (int i = 0; i < N; i++) c[i] = pos[i];
        Node<N> parent(g, c, parenti);
        parent.getNeigh(buckets.data(), vec_size);
        int count = 0;
        for (int b = 0; b < vec_size; b++) count += (int)buckets[b].size();
        put_i32(fo, parent.get_f());
        put_i32(fo, count);
        for (int b = 0; b < vec_size; b++) {
            for (size_t s = 0; s < buckets[b].size(); s++) {
                const Node<N> &nd = buckets[b][s];
                uint16_t p2[N];
                for (int i = 0; i < N; i++) p2[i] = nd.pos[i];
                fwrite(p2, 2, N, fo);
                put_i32(fo, nd.get_f());
                put_i32(fo, nd.get_g());
                put_i32(fo, nd.get_parenti());
                put_i32(fo, (int32_t)nd.pos.get_id(vec_size));
            }
            buckets[b].clear();
        }
    }
    fclose(fi);
    fclose(fo);
    return 0;
}

// ---------------------------------------------------------------------------
// owner: Coord<N>::get_id over coords read from a file.
//   in : i32 K, K x u16 pos[N];   out: "PGO1", i32 K, K x u32 owner
// ---------------------------------------------------------------------------
template <int N>
int owner_run(const char *in, const char *out, int vec_size)
{
    FILE *fi = fopen(in, "rb");
    FILE *fo = fopen(out, "wb");
    if (!fi || !fo) return 3;
    int32_t K = 0;
    if (fread(&K, 4, 1, fi) != 1) return 3;
    fwrite("PGO1", 4, 1, fo);
    put_i32(fo, K);
    for (int k = 0; k < K; k++) {
        uint16_t pos[N];
        if (fread(pos, 2, N, fi) != (size_t)N) return 3;
        Coord<N> c;
        for (int i = 0; i < N; i++) c[i] = pos[i];
        put_i32(fo, (int32_t)c.get_id(vec_size));
    }
    fclose(fi);
    fclose(fo);
    return 0;
}

// ---------------------------------------------------------------------------
// Open list with PriorityList.h semantics, std containers only.
// ---------------------------------------------------------------------------
template <int N>
class OpenListStd
{
  public:
    typedef std::multimap<int, Node<N> > ByF;
    bool dequeue(Node<N> &n) // PriorityList.h:84-93
    {
        typename ByF::iterator it = by_f.begin();
        if (it == by_f.end()) return false;
        n = it->second;
        by_pos.erase(n.pos);
        by_f.erase(it);
        return true;
    }
    void conditional_enqueue(const Node<N> &c) // PriorityList.h:104-113
    {
        typename std::map<Coord<N>, typename ByF::iterator>::iterator it = by_pos.find(c.pos);
        if (it == by_pos.end()) {
            by_pos[c.pos] = by_f.insert(std::make_pair(c.get_f(), c));
            return;
        }
        if (c.get_f() >= it->second->second.get_f()) return;
        by_f.erase(it->second);
        it->second = by_f.insert(std::make_pair(c.get_f(), c));
    }
    int get_highest_priority() const // PriorityList.h:115-122
    {
        if (by_f.empty()) return std::numeric_limits<int>::max();
        return by_f.begin()->first;
    }
    size_t size() const { return by_f.size(); }
    bool empty() const { return by_f.empty(); }

  private:
    ByF by_f;
    std::map<Coord<N>, typename ByF::iterator> by_pos;
};

struct SearchStats {
    int g_final = -1;
    int f_final = -1;
    long long pops = 0;       // the reference's "Total" (PAStar.cpp:341)
    long long expansions = 0; // passed the closed check and not the goal
    long long generated = 0;
    long long reopen = 0;
    long long open_size = 0;
    long long closed_size = 0;
    double seconds = 0;
    int finished = 0;
    // measured after the first `warm_pops` dequeues (bench.py --impl reference: untimed warm-up, then K steps)
    double timed_seconds = 0;
    long long timed_expansions = 0;
};

// Walk parenti from the goal (backtrace.cpp:44-69) and emit the aligned rows.
template <int N>
void backtrace_rows(std::map<Coord<N>, Node<N> > *closed, int lists, std::vector<std::string> &rows)
{
    Sequences *seq = Sequences::getInstance();
    Coord<N> fin = Sequences::get_final_coord<N>();
    rows.assign(N, std::string());
    Node<N> cur = closed[fin.get_id(lists)][fin];
    do {
        for (int i = 0; i < N; i++) {
            char c = '-';
            if (cur.pos[i] != cur.get_parent()[i]) c = seq->get_seq(i)[cur.pos[i] - 1];
            rows[i].insert(rows[i].begin(), c);
        }
        Coord<N> par = cur.get_parent();
        cur = closed[par.get_id(lists)][par];
    } while (cur.pos != Sequences::get_initial_coord<N>());
}

// Serial A*, AStar.cpp:53-104.  `budget` > 0 stops after that many pops.
template <int N>
SearchStats astar_run(long long budget, std::vector<std::string> *rows)
{
    SearchStats st;
    Node<N> current;
    OpenListStd<N> open;
    std::map<Coord<N>, Node<N> > closed;
    std::vector<Node<N> > neigh;
    Coord<N> coord_final = Sequences::get_final_coord<N>();
    Coord<N>::configure_hash(HashFZorder, 12);
    double t0 = now_s();
    open.conditional_enqueue(Sequences::get_initial_node<N>());
    while (!open.empty()) {
        typename std::map<Coord<N>, Node<N> >::iterator c_search;
        open.dequeue(current);
        st.pops++;
        if ((c_search = closed.find(current.pos)) != closed.end()) {
            if (current.get_g() >= c_search->second.get_g()) continue;
            st.reopen++;
        }
        closed[current.pos] = current;
        if (current.pos == coord_final) {
            st.g_final = current.get_g();
            st.f_final = current.get_f();
            st.finished = 1;
            break;
        }
        st.expansions++;
        current.getNeigh(&neigh);
        st.generated += (long long)neigh.size();
        for (typename std::vector<Node<N> >::iterator it = neigh.begin(); it != neigh.end(); ++it) {
            if ((c_search = closed.find(it->pos)) != closed.end()) {
                if (it->get_g() >= c_search->second.get_g()) continue;
                closed.erase(it->pos);
            }
            open.conditional_enqueue(*it);
        }
        neigh.clear();
        if (budget > 0 && st.pops >= budget) break;
    }
    st.seconds = now_s() - t0;
    st.open_size = (long long)open.size();
    st.closed_size = (long long)closed.size();
    if (st.finished && rows) backtrace_rows<N>(&closed, 1, *rows);
    return st;
}

// ---------------------------------------------------------------------------
// T-thread hash-partitioned PA-Star, one process.  Follows PAStar.cpp.
// ---------------------------------------------------------------------------
template <int N>
class PAStarStd
{
  public:
    PAStarStd(int threads, long long budget, long long warm_pops = 0)
        : T(threads), budget(budget), warm_pops(warm_pops), open(threads), closed(threads), inbox(threads), inbox_mutex(threads),
          inbox_cv(threads), pops(threads, 0), expansions(threads, 0), generated(threads, 0), reopen(threads, 0)
    {
        end_cond = false;
        end_condLocal = false;
        sync_count = 0;
        final_node.set_max();
        final_node_count = 0;
        total_pops = 0;
        total_expansions = 0;
        warm_time = 0;
        warm_expansions = 0;
        budget_hit = false;
        // PAStar.cpp:153: the start node goes to OpenList[0] whatever its owner
        open[0].conditional_enqueue(Sequences::get_initial_node<N>());
    }

    SearchStats run(std::vector<std::string> *rows)
    {
        SearchStats st;
        Coord<N> coord_final = Sequences::get_final_coord<N>();
        double t0 = now_s();
        std::vector<std::thread> th;
        for (int i = 0; i < T; i++) th.push_back(std::thread(&PAStarStd::worker, this, i, coord_final));
        for (size_t i = 0; i < th.size(); i++) th[i].join();
        const double t1 = now_s();
        st.seconds = t1 - t0;
        for (int i = 0; i < T; i++) {
            st.pops += pops[i];
            st.expansions += expansions[i];
            st.generated += generated[i];
            st.reopen += reopen[i];
            st.open_size += (long long)open[i].size();
            st.closed_size += (long long)closed[i].size();
        }
        st.timed_seconds = t1 - (warm_pops > 0 && warm_time > 0 ? warm_time : t0);
        st.timed_expansions = st.expansions - (warm_pops > 0 && warm_time > 0 ? (long long)warm_expansions : 0);
        if (!budget_hit) {
            st.finished = 1;
            st.g_final = final_node.get_g();
            st.f_final = final_node.get_f();
            if (rows) backtrace_rows<N>(closed.data(), T, *rows);
        }
        return st;
    }

  private:
    int T;
    long long budget, warm_pops;
    double warm_time;
    std::atomic<long long> total_expansions, warm_expansions;
    std::vector<OpenListStd<N> > open;
    std::vector<std::map<Coord<N>, Node<N> > > closed;
    std::vector<std::vector<Node<N> > > inbox;
    std::vector<std::mutex> inbox_mutex;
    std::vector<std::condition_variable> inbox_cv;
    std::vector<long long> pops, expansions, generated, reopen;
    std::atomic<bool> end_cond, end_condLocal, budget_hit;
    std::atomic<long long> total_pops;
    std::mutex final_node_mutex;
    Node<N> final_node;
    std::atomic<int> final_node_count;
    std::mutex sync_mutex;
    std::condition_variable sync_cv;
    int sync_count;

    void enqueue(int tid, std::vector<Node<N> > &nodes) // PAStar.cpp:219-237
    {
        typename std::map<Coord<N>, Node<N> >::iterator c_search;
        for (typename std::vector<Node<N> >::iterator it = nodes.begin(); it != nodes.end(); ++it) {
            if ((c_search = closed[tid].find(it->pos)) != closed[tid].end()) {
                if (it->get_g() >= c_search->second.get_g()) continue;
                closed[tid].erase(it->pos);
                reopen[tid] += 1;
            }
            open[tid].conditional_enqueue(*it);
        }
    }
    void consume_queue(int tid) // PAStar.cpp:240-250
    {
        std::unique_lock<std::mutex> lk(inbox_mutex[tid]);
        std::vector<Node<N> > nodes(inbox[tid]);
        inbox[tid].clear();
        lk.unlock();
        enqueue(tid, nodes);
    }
    void wait_queue(int tid) // PAStar.cpp:253-260 (bounded wait so a budget stop cannot hang)
    {
        std::unique_lock<std::mutex> lk(inbox_mutex[tid]);
        if (inbox[tid].size() == 0) inbox_cv[tid].wait_for(lk, std::chrono::milliseconds(1));
    }
    void wake_all_queue()
    {
        for (int i = 0; i < T; i++) {
            std::unique_lock<std::mutex> lk(inbox_mutex[i]);
            inbox_cv[i].notify_one();
        }
    }
    void sync_threads() // PAStar.cpp:276-316 without the MPI barriers
    {
        std::unique_lock<std::mutex> lk(sync_mutex);
        if (++sync_count < T) {
            sync_cv.wait(lk);
        } else {
            sync_count = 0;
            sync_cv.notify_all();
        }
    }
    void process_final_node(int tid, const Node<N> &n) // PAStar.cpp:410-468
    {
        std::unique_lock<std::mutex> lk(final_node_mutex);
        if (final_node.get_f() < n.get_f()) return;
        if (n.pos.get_id(T) == (unsigned int)tid) {
            final_node = n;
            final_node_count = 0;
            for (int i = 0; i < T; i++) {
                if (i != tid) {
                    std::lock_guard<std::mutex> ql(inbox_mutex[i]);
                    inbox[i].push_back(n);
                    inbox_cv[i].notify_one();
                }
            }
        }
        lk.unlock();
        if (++final_node_count == T) end_condLocal = true;
    }
    bool check_stop(int tid) // PAStar.cpp:479-547, single rank
    {
        Node<N> n = final_node;
        wake_all_queue();
        sync_threads();
        if (budget_hit) return false;
        consume_queue(tid);
        if (open[tid].get_highest_priority() < final_node.get_f()) end_condLocal = false;
        sync_threads();
        if (tid == 0) end_cond = (bool)end_condLocal;
        sync_threads();
        if (!end_cond) {
            if (!end_condLocal) {
                closed[tid].erase(n.pos);
                if (n.pos.get_id(T) == (unsigned int)tid) open[tid].conditional_enqueue(n);
            }
            return true;
        }
        return false;
    }
    void worker_inner(int tid, const Coord<N> &coord_final) // PAStar.cpp:319-401
    {
        Node<N> current;
        std::vector<std::vector<Node<N> > > neigh(T);
        while (end_condLocal == false) {
            typename std::map<Coord<N>, Node<N> >::iterator c_search;
            if (budget > 0 && total_pops >= budget) {
                budget_hit = true;
                end_condLocal = true;
                break;
            }
            consume_queue(tid);
            if (open[tid].dequeue(current) == false) {
                wait_queue(tid);
                continue;
            }
            pops[tid] += 1;
            if (++total_pops == warm_pops) {
                warm_expansions = (long long)total_expansions;
                warm_time = now_s();
            }
            if ((c_search = closed[tid].find(current.pos)) != closed[tid].end()) {
                if (current.get_g() >= c_search->second.get_g()) continue;
                reopen[tid] += 1;
            }
            closed[tid][current.pos] = current;
            if (current.pos == coord_final) {
                process_final_node(tid, current);
                continue;
            }
            expansions[tid] += 1;
            total_expansions++;
            current.getNeigh(neigh.data(), T);
            for (int i = 0; i < T; i++) {
                generated[tid] += (long long)neigh[i].size();
                if (i == tid) {
                    enqueue(tid, neigh[i]);
                } else if (neigh[i].size() != 0) {
                    std::unique_lock<std::mutex> lk(inbox_mutex[i]);
                    inbox[i].insert(inbox[i].end(), neigh[i].begin(), neigh[i].end());
                    lk.unlock();
                    inbox_cv[i].notify_one();
                }
                neigh[i].clear();
            }
        }
    }
    void worker(int tid, Coord<N> coord_final) // PAStar.cpp:550-589 (no affinity: F10)
    {
        sync_threads();
        do {
            worker_inner(tid, coord_final);
        } while (check_stop(tid));
    }
};

void print_stats(const char *what, int n, int threads, const SearchStats &st, const std::vector<std::string> &rows)
{
    printf("{\"cmd\": \"%s\", \"n_seq\": %d, \"threads\": %d, \"finished\": %d, \"g\": %d, \"f\": %d, "
           "\"pops\": %lld, \"expansions\": %lld, \"generated\": %lld, \"reopen\": %lld, "
           "\"open\": %lld, \"closed\": %lld, \"seconds\": %.6f, \"timed_seconds\": %.6f, \"timed_expansions\": %lld",
           what, n, threads, st.finished, st.g_final, st.f_final, st.pops, st.expansions, st.generated, st.reopen,
           st.open_size, st.closed_size, st.seconds, st.timed_seconds, st.timed_expansions);
    if (!rows.empty()) {
        printf(", \"rows\": [");
        for (size_t i = 0; i < rows.size(); i++) printf("%s\"%s\"", i ? ", " : "", rows[i].c_str());
        printf("]");
    }
    printf("}\n");
}

template <int N>
int search_run(const std::string &cmd, int threads, long long budget, hashType ht, int shift, long long warm)
{
    std::vector<std::string> rows;
    SearchStats st;
    if (cmd == "astar") {
        st = astar_run<N>(budget, &rows);
    } else {
        Coord<N>::configure_hash(ht, shift);
        PAStarStd<N> p(threads, budget, warm);
        st = p.run(&rows);
    }
    print_stats(cmd.c_str(), N, threads, st, rows);
    return 0;
}

// Micro-baselines: PairAlign GCUPS and getNeigh-only parents/s on one core.
template <int N>
int micro_run(double seconds)
{
    Sequences *seq = Sequences::getInstance();
    long long cells = 0;
    for (int i = 0; i < N - 1; i++)
        for (int j = i + 1; j < N; j++)
            cells += (long long)(seq->get_seq(i).length() + 1) * (long long)(seq->get_seq(j).length() + 1);
    double t0 = now_s();
    int reps = 0;
    do {
        std::vector<PairAlign *> v = build_pairs();
        for (size_t i = 0; i < v.size(); i++) delete v[i];
        reps++;
    } while (now_s() - t0 < seconds * 0.5);
    double dp_s = (now_s() - t0) / reps;

    Coord<N> fin = Sequences::get_final_coord<N>();
    std::vector<Node<N> > parents;
    uint64_t s = 88172645463325252ull;
    for (int k = 0; k < 4096; k++) {
        Coord<N> c;
        for (int i = 0; i < N; i++) {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            c[i] = (uint16_t)(s % (uint64_t)(fin[i] > 1 ? fin[i] - 1 : 1));
        }
        parents.push_back(Node<N>(1000, c, (1 << N) - 1));
    }
    std::vector<Node<N> > out;
    long long done = 0, succ = 0;
    t0 = now_s();
    do {
        for (size_t k = 0; k < parents.size(); k++) {
            parents[k].getNeigh(&out);
            succ += (long long)out.size();
            out.clear();
        }
        done += (long long)parents.size();
    } while (now_s() - t0 < seconds * 0.5);
    double ne_s = now_s() - t0;
    printf("{\"cmd\": \"micro\", \"n_seq\": %d, \"dp_cells\": %lld, \"dp_seconds\": %.6f, \"dp_gcups\": %.6f, "
           "\"neigh_parents\": %lld, \"neigh_successors\": %lld, \"neigh_seconds\": %.6f, "
           "\"neigh_parents_per_s\": %.1f}\n",
           N, cells, dp_s, cells / dp_s * 1e-9, done, succ, ne_s, done / ne_s);
    return 0;
}

// The reference's own printer (backtrace.cpp:135-191, through print_entire_backtrace as PAStarDistributedBacktrace.cpp:211
// calls it) over aligned rows read from a text file: N lines of equal length.
template <int N>
int print_run(const char *rows_path)
{
    std::ifstream in(rows_path);
    std::list<char> alignments[N];
    for (int i = 0; i < N; i++) {
        std::string line;
        if (!std::getline(in, line)) return 1;
        alignments[i].assign(line.begin(), line.end());
    }
    print_entire_backtrace<N>(alignments);
    return 0;
}

int usage()
{
    std::cerr << "usage: pastar_ref dump  <fasta> <out.bin>\n"
                 "       pastar_ref neigh <fasta> <parents.bin> <out.bin> <vec_size> <HASH> <shift>\n"
                 "       pastar_ref owner <fasta> <coords.bin> <out.bin> <vec_size> <HASH> <shift>\n"
                 "       pastar_ref astar <fasta> [budget]\n"
                 "       pastar_ref pastar <fasta> <threads> [budget] [HASH] [shift] [warm_pops]\n"
                 "       pastar_ref micro <fasta> [seconds]\n"
                 "       pastar_ref seqs  <fasta>\n"
                 "       pastar_ref print <fasta> <rows.txt>\n";
    return 2;
}

} // namespace

int main(int argc, char **argv)
{
    if (argc < 3) return usage();
    std::string cmd = argv[1];
    if (read_fasta_file(argv[2]) != 0) return 1;
    int n = Sequences::get_seq_num();
    if (cmd == "seqs") { // what read_fasta_file (read_fasta.cpp:8-56) made of the file: count, then one sequence per line
        std::cout << n << "\n";
        for (int i = 0; i < n; i++) std::cout << Sequences::getInstance()->get_seq(i) << "\n";
        std::cout.flush();
        _exit(0); // the heuristic was never initialised: skip the singletons' destructors
    }
    if (cmd == "print") {
        if (argc < 4) return usage();
#define DISPATCH_PRINT_RC(X)         \
    case X:                          \
        rc = print_run<X>(argv[3]); \
        break;
        int rc = 1;
        switch (n) { MAX_NUM_SEQ_HELPER(DISPATCH_PRINT_RC) }
        std::cout.flush();
        _exit(rc);
    }
    // stdout of init ("Starting pairwise alignments..." + timer) is noise for
    // the JSON consumers: send it to stderr.
    std::streambuf *keep = std::cout.rdbuf(std::cerr.rdbuf());
    double t0 = now_s();
    HeuristicHPair::getInstance()->init();
    double init_s = now_s() - t0;
    std::cout.rdbuf(keep);
    (void)init_s;

#define DISPATCH(X) \
    case X:         \
        return CALL(X);
    if (cmd == "dump") {
        if (argc < 4) return usage();
        return cmd_dump(argv[3]);
    } else if (cmd == "neigh" || cmd == "owner") {
        if (argc < 8) return usage();
        int vec = atoi(argv[5]);
        // the hash is a process-global in the reference (CoordHash.cpp:17-18)
        Coord<3>::configure_hash(parse_hash(argv[6]), atoi(argv[7]));
        if (cmd == "neigh") {
#define CALL(X) neigh_run<X>(argv[3], argv[4], vec)
            switch (n) { MAX_NUM_SEQ_HELPER(DISPATCH) }
#undef CALL
        } else {
#define CALL(X) owner_run<X>(argv[3], argv[4], vec)
            switch (n) { MAX_NUM_SEQ_HELPER(DISPATCH) }
#undef CALL
        }
    } else if (cmd == "astar" || cmd == "pastar") {
        int threads = 1;
        long long budget = 0;
        hashType ht = HashFZorder;
        int shift = 12;
        long long warm = 0;
        if (cmd == "astar") {
            if (argc > 3) budget = atoll(argv[3]);
        } else {
            if (argc < 4) return usage();
            threads = atoi(argv[3]);
            if (argc > 4) budget = atoll(argv[4]);
            if (argc > 5) ht = parse_hash(argv[5]);
            if (argc > 6) shift = atoi(argv[6]);
            if (argc > 7) warm = atoll(argv[7]);
        }
#define CALL(X) search_run<X>(cmd, threads, budget, ht, shift, warm)
        switch (n) { MAX_NUM_SEQ_HELPER(DISPATCH) }
#undef CALL
    } else if (cmd == "micro") {
        double secs = argc > 3 ? atof(argv[3]) : 2.0;
#define CALL(X) micro_run<X>(secs)
        switch (n) { MAX_NUM_SEQ_HELPER(DISPATCH) }
#undef CALL
    } else {
        return usage();
    }
    std::cerr << "unsupported number of sequences: " << n << std::endl;
    return 1;
}
