// Binds the drop-in boundary from C++, the way a maintainer of the reference would: the host mirror's classes keep the
// reference's names and signatures (pastar/include/Node.h:37, Coord.h:49, HeuristicHPair.h:18-23, Sequences.h:19-28),
// behind them sits the C ABI of include/pastar_gpu.h.  Prints, for a FASTA file and a list of parent nodes on stdin
// ("c0 c1 .. cN-1 g parenti" per line), what the reference's own objects would give:
//   H <calculate_h>
//   ID <get_id(vec_size)>
//   S <owner> <pos..> <f> <g> <parenti>      one line per successor of Node::getNeigh, bucket by bucket
// tests/test_cpp_boundary.py diffs the output with the oracle (pinned to the reference compiled in place) and with
// oracle/_ref/pastar_ref's own dump when the reference sources are present.
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "host/pastar_host.hpp"

#include <fstream>

namespace pastar {
int hash_shift = HASH_SHIFT;      // CoordHash.cpp:17 (the CLI defines these too; this program stands in for it)
hashType hash_type = HashFZorder; // CoordHash.cpp:18
// read_fasta_file_core, pastar/read_fasta.cpp:8-36: '>' lines and empty lines end a record
int read_fasta_file(const std::string &name)
{
    std::ifstream file(name.c_str());
    if (!file.is_open()) return -1;
    std::string seq, buf;
    while (std::getline(file, buf)) {
        if (buf.empty() || buf[0] == '>') {
            if (!seq.empty()) Sequences::getInstance()->set_seq(seq);
            seq.clear();
        } else {
            seq.append(buf);
        }
    }
    if (!seq.empty()) Sequences::getInstance()->set_seq(seq);
    return 0;
}
} // namespace pastar
using namespace pastar;

template <int N>
static int run(int vec_size)
{
    HeuristicHPair *h = HeuristicHPair::getInstance();
    std::string line;
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::istringstream in(line);
        Coord<N> c;
        for (int i = 0; i < N; i++) {
            int v;
            in >> v;
            c[i] = (uint16_t)v;
        }
        int g = 0, parenti = 0;
        in >> g >> parenti;
        std::cout << "H " << h->calculate_h<N>(c) << "\n";
        std::cout << "ID " << c.get_id(vec_size) << "\n";
        Node<N> node(g, c, parenti); // the constructor computes f = g + h (Node.cpp:32-39)
        std::vector<std::vector<Node<N>>> buckets(vec_size);
        node.getNeigh(buckets.data(), vec_size);
        for (int o = 0; o < vec_size; o++)
            for (const Node<N> &s : buckets[o]) {
                std::cout << "S " << o;
                for (int i = 0; i < N; i++) std::cout << " " << s.pos[i];
                std::cout << " " << s.get_f() << " " << s.get_g() << " " << s.get_parenti() << "\n";
            }
        std::cout << "E\n";
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::cerr << "usage: boundary_test file.fasta vec_size {FZORDER|PZORDER|FSUM|PSUM} shift < parents\n";
        return 2;
    }
    try {
        if (read_fasta_file(argv[1]) != 0) return 1;
        const int vec_size = std::stoi(argv[2]);
        const std::string ht = argv[3];
        const hashType type = ht == "FZORDER" ? HashFZorder : ht == "PZORDER" ? HashPZorder : ht == "FSUM" ? HashFSum : HashPSum;
        HeuristicHPair::getInstance()->init();
        const int n = Sequences::get_seq_num();
        switch (n) {
#define CASE(X)                                              \
    case X:                                                  \
        Coord<X>::configure_hash(type, std::stoi(argv[4])); \
        return run<X>(vec_size);
            CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(14) CASE(16)
#undef CASE
        default:
            std::cerr << "unsupported number of sequences\n";
            return 1;
        }
    } catch (const std::exception &e) {
        std::cerr << "Running fatal error: " << e.what() << std::endl;
        return -1;
    }
}
