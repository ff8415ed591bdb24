#!/bin/bash
# CPU only: the checker and the product's host-side code under UBSan / ASan (no GPU needed; nothing here launches a kernel).
#   1. oracle/pastar_oracle.c rebuilt with -fsanitize=undefined, its CPU tests run against that build, the normal build restored
#   2. host/pg_host_weights.cpp with ASan + UBSan over 60 seeded inputs (N 2..16, lengths 1..400, identical sequences)
#   3. the CLI's translation unit (option parser, FASTA reader, result printer) with ASan + UBSan
# Last run (end of round 2): all three clean.
set -e
R=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
trap 'rm -rf "$T"' EXIT
SAN="-O1 -g -fsanitize=undefined -fno-sanitize-recover=undefined"

cp "$R/oracle/libpastar_oracle.so" "$T/liborig.so"
gcc -std=c11 $SAN -fsanitize=bounds-strict -fPIC -shared -ffp-contract=off -o "$R/oracle/libpastar_oracle.so" "$R/oracle/pastar_oracle.c" -lm
(cd "$R" && python -m pytest tests/test_oracle_golden.py tests/test_oracle_vs_ref.py -x -q) || { cp "$T/liborig.so" "$R/oracle/libpastar_oracle.so"; exit 1; }
cp "$T/liborig.so" "$R/oracle/libpastar_oracle.so"

cat > "$T/w.cpp" <<'CPP'
#include "pastar_gpu.h"
#include <cstdio>
#include <string>
#include <vector>
int main()
{
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    const char *al = "ACDEFGHIKLMNPQRSTVWY";
    int bad = 0;
    for (int t = 0; t < 60; t++) {
        int n = 2 + (int)(rnd() % 15);
        std::vector<std::string> q(n);
        for (auto &x : q) {
            int L = 1 + (int)(rnd() % (t % 7 == 0 ? 400 : 40));
            for (int i = 0; i < L; i++) x.push_back(t % 5 == 0 ? "AC"[rnd() % 2] : al[rnd() % 20]);
        }
        if (t % 11 == 0) for (auto &x : q) x = q[0];
        std::vector<const char *> p(n);
        std::vector<int> len(n);
        for (int i = 0; i < n; i++) { p[i] = q[i].data(); len[i] = (int)q[i].size(); }
        std::vector<float> w((size_t)n * n);
        if (pg_host_weights(n, p.data(), len.data(), w.data())) bad++;
    }
    printf("pg_host_weights under ASan + UBSan: %d failures\n", bad);
    return bad;
}
CPP
INC="-I$R/include -I$R/mpi_pastar_msa_b200/csrc"
g++ -std=c++17 $SAN -fsanitize=address -ffp-contract=off $INC "$T/w.cpp" "$R/mpi_pastar_msa_b200/csrc/host/pg_host_weights.cpp" -pthread -o "$T/w"
"$T/w"

LIBDIR="$R/mpi_pastar_msa_b200/lib"
g++ -std=c++17 $SAN -fsanitize=address $INC -Dmain=pastar_cli_main -c "$R/mpi_pastar_msa_b200/csrc/host/pastar_main.cpp" -o "$T/m.o"
g++ -std=c++17 $SAN -fsanitize=address $INC "$R/tests/cpp/host_cpu_test.cpp" "$T/m.o" -L"$LIBDIR" -lpastar_gpu -Wl,-rpath,"$LIBDIR" -pthread -o "$T/h"
g++ -std=c++17 $SAN -fsanitize=address $INC "$R/mpi_pastar_msa_b200/csrc/host/pastar_main.cpp" -L"$LIBDIR" -lpastar_gpu -Wl,-rpath,"$LIBDIR" -pthread -o "$T/cli"
export ASAN_OPTIONS=detect_leaks=0
printf ">a\nACD\n\n>b\n>c\nAC\n>d\nGG" > "$T/f.fa"
printf "AC-D\nACED\nA--D\n" > "$T/rows.txt"
"$T/h" seqs "$T/f.fa" > /dev/null
"$T/h" print "$T/rows.txt" > /dev/null
for a in "--thr 4 $T/f.fa" "-t" "--hash_type=X $T/f.fa" "-h" "$T/f.fa -s 3 -yPSUM --batch=4" "--metrics_json" "-g2 $T/f.fa" ""; do
    "$T/cli" $a > /dev/null 2> "$T/e.txt" || true
    if grep -q "runtime error\|AddressSanitizer" "$T/e.txt"; then echo "sanitizer report for: pastar $a"; cat "$T/e.txt"; exit 1; fi
done
echo "CLI parser / FASTA reader / printer under ASan + UBSan: clean"
