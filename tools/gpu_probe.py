"""Quick on-GPU numbers while developing (not the bench): python tools/gpu_probe.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mpi_pastar_msa_b200 as m
from conftest import CASES, S7, S8, random_parents

def dp(name, seqs):
    with m.PastarGPU(seqs, weights=None) as G:
        ts = [G.build_pair_tables() for _ in range(5)]
        cells = sum((len(a)+1)*(len(b)+1) for i,a in enumerate(seqs) for b in seqs[i+1:])
        print("DP %s: cells %d best %.3f ms -> %.1f GCUPS (all: %s)" % (name, cells, min(ts), cells/min(ts)/1e6, ["%.3f"%t for t in ts]), flush=True)

def search(name, seqs, **kw):
    with m.PastarGPU(seqs) as G:
        G.build_pair_tables()
        t=time.time(); r = G.search(want_rows=False, **kw); dt=time.time()-t
        print("SEARCH %s %s: finished %d g %d exp %d gen %d pops %d reopen %d open %d closed %d rounds %d wall %.3fs dev %.1f ms -> %.2f Mexp/s %.2f Gsucc/s" % (
            name, kw, r["finished"], r["g"], r["expansions"], r["generated"], r["pops"], r["reopen"], r["open_size"], r["closed_size"], r["rounds"], dt, r["kernel_ms"],
            r["expansions"]/r["kernel_ms"]/1e3, r["generated"]/r["kernel_ms"]/1e6), flush=True)

def expand(name, seqs, K):
    import torch
    with m.PastarGPU(seqs) as G:
        G.build_pair_tables()
        n = G.n
        pos, g, par = random_parents(seqs, K, 1)
        lens = np.array(G.lens)
        pos = np.minimum(pos, lens - 1).astype(np.uint16)  # interior parents
        nodes = G.make_nodes(pos, g, par)
        d_par = torch.from_numpy(nodes.view(np.uint8).reshape(K, -1)).cuda()
        sst = m.succ_dtype(n).itemsize
        d_out = torch.empty(K * G.S * sst, dtype=torch.uint8, device="cuda")
        d_cnt = torch.empty(K, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for vs in (1, 8):
            for _ in range(3):
                G.expand_batch_dev(d_par.data_ptr(), K, vs, d_out.data_ptr(), d_cnt.data_ptr(), st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 5
            for _ in range(reps):
                G.expand_batch_dev(d_par.data_ptr(), K, vs, d_out.data_ptr(), d_cnt.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nst = m.node_dtype(n).itemsize
            bexp = nst + n + 16 * G.npairs + G.S * sst
            print("EXPAND %s K=%d vec=%d: %.3f ms -> %.1f Mexp/s %.1f Gsucc/s, %.0f GB/s algorithmic (%.0f B/exp)" % (
                name, K, vs, ms, K/ms/1e3, K*G.S/ms/1e6, K*bexp/ms/1e6, bexp), flush=True)

which = sys.argv[1:] or ["dp", "kinase", "s7", "expand"]
if "dp" in which:
    dp("kinase", CASES["kinase"]); dp("S7", S7()); dp("S8", S8())
if "kinase" in which:
    for bt in (1024, 4096, 16384, 65536):
        search("kinase", CASES["kinase"], table_capacity=1<<27, batch_target=bt)
if "s7" in which:
    for bt in (16384, 65536, 262144):
        search("S7", S7(), table_capacity=1<<28, batch_target=bt, max_expansions=3000000)
if "expand" in which:
    expand("S7", S7(), 200000)
    expand("S8", S8()[:8] if False else [s[:998] for s in S8()], 100000)
