"""One run of the stand-alone getNeigh-batch kernel for profiling: python tools/expand_once.py [s7|s8] [parents]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mpi_pastar_msa_b200 as m
from conftest import S7, S8
which = sys.argv[1] if len(sys.argv) > 1 else "s7"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
seqs = S8() if which == "s8" else S7()
n, L = len(seqs), len(seqs[0])
with m.PastarGPU(seqs, weights=None) as G:
    G.build_pair_tables()
    rng = np.random.default_rng(1)
    pos = np.stack([rng.integers(0, L, K) for _ in range(n)], axis=1).astype(np.uint16)
    nodes = G.make_nodes(pos, rng.integers(0, 100000, K), rng.integers(1, 1 << n, K))
    d_par = torch.from_numpy(nodes.view(np.uint8).reshape(K, -1)).cuda()
    d_out = torch.empty(K * G.S * m.succ_dtype(n).itemsize, dtype=torch.uint8, device="cuda")
    d_cnt = torch.empty(K, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(4):
        G.expand_batch_dev(d_par.data_ptr(), K, 8, d_out.data_ptr(), d_cnt.data_ptr(), st)
    torch.cuda.synchronize()
    print("ok", int(d_cnt.sum().item()))
