"""torchrun --nproc-per-node G tools/multi_gpu_check.py : partitioned search parity on G GPUs (optimal cost + alignment)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import mpi_pastar_msa_b200 as m
from mpi_pastar_msa_b200.dist import CudaEngine, CudaEngineP2P, PartitionedSearch
P2P = os.environ.get('PG_P2P', '1') == '1'
FWD = os.environ.get('PG_FWD', '1') == '1'
from conftest import CASES, KNOWN_OPT, weighted_sp_score
from oracle import oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, batch, ht, sh in [("PF08184", 64, "FZORDER", 3), ("test2", 256, "FSUM", 1), ("fam5x60", 1024, "FZORDER", 2), ("fam8x20", 4096, "PZORDER", 1),
                            ("fam4x150", 4096, "FZORDER", 12), ("kinase", 16384, "FZORDER", 12), ("kinase", 65536, "FZORDER", 4)]:
    seqs = CASES[name]
    G = m.PastarGPU(seqs, device=local)
    G.build_pair_tables()
    G.configure_hash(ht, sh)
    eng = CudaEngineP2P(G, world, rank, dist, 1 << 26, batch, forward=FWD) if P2P else CudaEngine(G, world, rank, 1 << 26, batch)
    drv = PartitionedSearch(eng, dist, seqs, lambda pos: int(G.owner(np.array(pos, dtype=np.uint16), world)[0]))
    torch.cuda.synchronize(); t0 = time.time()
    r = drv.run()
    torch.cuda.synchronize(); dt = time.time() - t0
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    good = r["finished"] == 1 and r["g"] == ref and weighted_sp_score(seqs, G.w_int, r["rows"]) == ref
    ok = ok and good
    if rank == 0:
        print("%s %-9s batch %6d %s/%d: g %d (ref %d) exp %d gen %d rounds %d  %.3fs  sent %.1f MB/rank" % (
            "OK " if good else "BAD", name, batch, ht, sh, r["g"], ref, r["expansions"], r["generated"], r["rounds"], dt, drv.bytes_sent / 1e6), flush=True)
    eng.end(); G.close()
dist.barrier()
if rank == 0:
    print("ALL OK" if ok else "FAILURES")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
