import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
import mpi_pastar_msa_b200 as m
from conftest import S7
seqs = S7()
torch.cuda.init(); torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); w = m.host_weights(seqs); t1 = time.perf_counter()
    G = m.PastarGPU(seqs); t2 = time.perf_counter()
    G.build_pair_tables(); t3 = time.perf_counter()
    G.search_begin(1, 0, 1 << 30, 1 << 20); t4 = time.perf_counter()
    G.search_rounds(8); t5 = time.perf_counter()
    G.search_end(); t6 = time.perf_counter()
    G.close(); t7 = time.perf_counter()
    print("weights %.1f ms, ctx (incl. weights) %.1f, tables %.1f, search_begin %.1f, 8 rounds %.1f, search_end %.1f, close %.1f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6)))
