"""Pairwise-DP kernel time against sequence length (fixed cost vs per-super-step cost)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpi_pastar_msa_b200 as m
from conftest import random_seqs
for n, L in [(3, 8), (7, 8), (7, 64), (7, 128), (7, 256), (7, 500), (7, 1000), (3, 1000), (3, 2000)]:
    seqs = random_seqs(n, L, 5)
    with m.PastarGPU(seqs, weights=None) as G:
        t = min(G.build_pair_tables() for _ in range(5))
    cells = n * (n - 1) // 2 * (L + 1) ** 2
    print("n %d L %5d : %.4f ms  %.1f GCUPS" % (n, L, t, cells / t / 1e6), flush=True)
