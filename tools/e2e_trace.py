"""torchrun --nproc-per-node G tools/e2e_trace.py [rounds_per_status] [budget_M]: wall time of the S7 partitioned search per
status interval, from the start node (what bench.py's N > 1 e2e leg runs): where a job's time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import mpi_pastar_msa_b200 as m
from mpi_pastar_msa_b200.dist import CudaEngineP2P, PartitionedSearch
from conftest import S7
per = int(sys.argv[1]) if len(sys.argv) > 1 else 8
budget = int(float(sys.argv[2]) * 1e6) if len(sys.argv) > 2 else 300_000_000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seqs = S7()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
G = m.PastarGPU(seqs, device=local)
G.set_stream(torch.cuda.current_stream().cuda_stream)
G.build_pair_tables()
t1 = time.perf_counter()
G.configure_hash("PZORDER", 6)
eng = CudaEngineP2P(G, world, rank, dist, 1 << 30, 1 << 20, forward=True)
drv = PartitionedSearch(eng, dist, seqs, lambda pos: int(G.owner(np.array(pos, dtype=np.uint16), world)[0]))
t2 = time.perf_counter()
best, last_t, last_e, k = 2**31 - 1, t2, 0, 0
rows = []
while True:
    mn, best, tot = drv.step(best, per)
    k += 1
    now = time.perf_counter()
    if k % 10 == 0:
        rows.append((drv.rounds, (now - last_t) / (10 * per) * 1e3, (tot[0] - last_e) / (10 * per)))
        last_t, last_e = now, tot[0]
    if tot[0] >= budget or mn >= best:
        break
torch.cuda.synchronize(); dist.barrier()
t3 = time.perf_counter()
if rank == 0:
    print("context + tables %.3f s, engine set-up %.3f s, search %.3f s for %d expansions in %d rounds (%d rounds per status)" % (t1 - t0, t2 - t1, t3 - t2, tot[0], drv.rounds, per))
    for r, ms, ex in rows:
        print("  rounds ..%5d: %.3f ms per round, %8.0f expansions per round" % (r, ms, ex))
eng.end(); G.close()
dist.destroy_process_group()
