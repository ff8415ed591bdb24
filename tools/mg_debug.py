import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import mpi_pastar_msa_b200 as m
from mpi_pastar_msa_b200.dist import CudaEngine, PartitionedSearch
from conftest import CASES
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seqs = CASES["PF08184"]
G = m.PastarGPU(seqs, device=local); G.build_pair_tables(); G.configure_hash("FZORDER", 3)
eng = CudaEngine(G, world, rank, 1 << 22, 64)
drv = PartitionedSearch(eng, dist, seqs, None)
best = 2**31 - 1
for rnd in range(12):
    out = eng.round(best)
    rec = np.dtype([("key", "<u8"), ("f", "<u4"), ("g", "<u4"), ("mask", "<u8")])
    desc = []
    for d, b in enumerate(out):
        if b.numel():
            a = b.cpu().numpy().view(rec)
            desc.append((d, [(hex(int(x["key"])), int(x["g"]), int(x["f"]), int(x["mask"])) for x in a]))
    inbox = drv.exchange(out)
    got = inbox.cpu().numpy().view(rec)
    eng.insert(inbox)
    mn, bg, cnt = eng.status()
    print("rank %d round %d: sent %s | recv %s | min_open_f %d best %d exp %d pushed %d inserted %d" % (
        rank, rnd, desc, [(hex(int(x["key"])), int(x["g"]), int(x["f"]), int(x["mask"])) for x in got], mn, bg, cnt["expansions"], cnt["pushed"], cnt["inserted"]), flush=True)
    red = torch.tensor([mn, bg], dtype=torch.int64, device="cuda"); dist.all_reduce(red, op=dist.ReduceOp.MIN)
    best = int(red[1])
dist.destroy_process_group()
