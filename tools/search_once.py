"""One search run for profiling: python tools/search_once.py {kinase|s7} [batch] [max_expansions]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mpi_pastar_msa_b200 as m
from conftest import CASES, S7
which = sys.argv[1]
bt = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
seqs = CASES["kinase"] if which == "kinase" else S7()
with m.PastarGPU(seqs) as G:
    G.build_pair_tables()
    cap = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 27
    r = G.search(want_rows=False, table_capacity=cap, batch_target=bt, max_expansions=mx)
    print(r)
