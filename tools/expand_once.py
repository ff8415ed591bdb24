import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import mpi_pastar_msa_b200 as m
from conftest import S7, S8
which = sys.argv[1] if len(sys.argv) > 1 else "s7"
seqs = S7() if which == "s7" else [s[:998] for s in S8()]
K = 100000
with m.PastarGPU(seqs) as G:
    G.build_pair_tables()
    n = G.n
    rng = np.random.default_rng(1)
    pos = np.stack([rng.integers(0, len(seqs[0]) - 1, K) for _ in range(n)], axis=1).astype(np.uint16)
    nodes = G.make_nodes(pos, rng.integers(0, 100000, K), rng.integers(1, 1 << n, K))
    d_par = torch.from_numpy(nodes.view(np.uint8).reshape(K, -1)).cuda()
    sst = m.succ_dtype(n).itemsize
    d_out = torch.empty(K * G.S * sst, dtype=torch.uint8, device="cuda")
    d_cnt = torch.empty(K, dtype=torch.int32, device="cuda")
    for vs in (8,):
        for _ in range(3):
            G.expand_batch_dev(d_par.data_ptr(), K, vs, d_out.data_ptr(), d_cnt.data_ptr(), 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            G.expand_batch_dev(d_par.data_ptr(), K, vs, d_out.data_ptr(), d_cnt.data_ptr(), 0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        bexp = m.node_dtype(n).itemsize + n + 16 * G.npairs + G.S * sst
        print("EXPAND %s K=%d vec=%d: %.3f ms -> %.1f Mexp/s %.1f Gsucc/s, %.0f GB/s algorithmic" % (which, K, vs, ms, K/ms/1e3, K*G.S/ms/1e6, K*bexp/ms/1e6))
