// Maintainer-side check of the host mirror's CPU-only pieces (no device call is made):
//   host_cpu_test seqs  <fasta>             what read_fasta_file made of the file: count, then one sequence per line
//   host_cpu_test print <rows.txt>          "Similarity:" + the wrapped alignment blocks for the aligned rows in the file
// tests/test_cpp_host_cpu.py compares both with the reference's own read_fasta.cpp / backtrace.cpp (oracle/_ref).
#include "host/pastar_host.hpp"

int main(int argc, char *argv[])
{
    using namespace pastar;
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    if (cmd == "seqs") {
        if (read_fasta_file(argv[2]) != 0) return 1;
        const int n = Sequences::get_seq_num();
        std::cout << n << "\n";
        for (int i = 0; i < n; i++) std::cout << Sequences::getInstance()->get_seq(i) << "\n";
        return 0;
    }
    if (cmd == "print") {
        std::ifstream in(argv[2]);
        std::vector<std::string> rows;
        for (std::string line; std::getline(in, line);) rows.push_back(line);
        print_similarity(rows);
        print_alignment(rows);
        return 0;
    }
    return 2;
}
