"""torchrun --nproc-per-node G tools/mg_profile.py [batch]: per-phase wall time of the partitioned round (synchronised, debug only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import mpi_pastar_msa_b200 as m
from mpi_pastar_msa_b200.dist import CudaEngine, CudaEngineP2P, PartitionedSearch
P2P = os.environ.get('PG_P2P', '1') == '1'
from conftest import S7
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
seqs = S7()
G = m.PastarGPU(seqs, device=local); G.build_pair_tables(); G.configure_hash("FZORDER", 12)
eng = CudaEngineP2P(G, world, rank, dist, 1 << 30, batch) if P2P else CudaEngine(G, world, rank, 1 << 30, batch)
drv = PartitionedSearch(eng, dist, seqs, None)
for _ in range(330):
    drv.step()
T = {"round": 0, "exchange": 0, "insert": 0, "status": 0, "allreduce": 0}
sync = torch.cuda.synchronize
n = 20
c0 = eng.status()[2]
recs = 0
for _ in range(n):
    sync(); dist.barrier(); sync(); t = time.perf_counter()
    if P2P:
        eng.g.search_round(2**31 - 1); sync(); t1 = time.perf_counter(); T["round"] += t1 - t
        allc = torch.empty(world * world, dtype=torch.int64, device="cuda"); dist.all_gather_into_tensor(allc, eng.counts)
        mine = allc.view(world, world)[:, rank].tolist(); mine[rank] = 0; sync(); t2 = time.perf_counter(); T["exchange"] += t2 - t1
        recs += sum(mine)
        eng.g.search_insert_segments_dev(eng.inbox.data_ptr(), eng.region, mine); sync(); t3 = time.perf_counter(); T["insert"] += t3 - t2
    else:
        out = eng.round(2**31 - 1); sync(); t1 = time.perf_counter(); T["round"] += t1 - t
        inbox = drv.exchange(out); sync(); t2 = time.perf_counter(); T["exchange"] += t2 - t1
        recs += inbox.numel() // eng.xrec
        eng.insert(inbox); sync(); t3 = time.perf_counter(); T["insert"] += t3 - t2
    mn, bg, cnt = eng.status(); sync(); t4 = time.perf_counter(); T["status"] += t4 - t3
    red = torch.tensor([mn, bg], dtype=torch.int64, device="cuda"); dist.all_reduce(red, op=dist.ReduceOp.MIN)
    tot = torch.tensor([1, 2, 3], dtype=torch.int64, device="cuda"); dist.all_reduce(tot); sync(); T["allreduce"] += time.perf_counter() - t4
c1 = eng.status()[2]
print("rank %d batch %d: per round ms %s | expansions/round %d, received records/round %d, pushed/round %d" % (
    rank, batch, {k: round(1e3 * v / n, 3) for k, v in T.items()}, (c1["expansions"] - c0["expansions"]) // n, recs // n, (c1["pushed"] - c0["pushed"]) // n), flush=True)
dist.destroy_process_group()
