import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mpi_pastar_msa_b200 as m
from conftest import S7, S8
seqs = S8() if (len(sys.argv) > 1 and sys.argv[1] == "s8") else S7()
with m.PastarGPU(seqs, weights=None) as G:
    for _ in range(3):
        print(G.build_pair_tables())
