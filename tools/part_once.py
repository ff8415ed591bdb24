"""G hash-owned partitions of the S7 search on ONE device (peer-mapped inboxes are plain device buffers, the cross-GPU
barrier a device synchronise): per-kernel times of the multi-partition kernels (MODE 2 expand, forward, inbox) without
needing G GPUs.   python tools/part_once.py [parts] [shift] [rounds] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mpi_pastar_msa_b200 as m
from conftest import S7
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 17
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 450
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
htype = sys.argv[5] if len(sys.argv) > 5 else "FZORDER"
cap = int(sys.argv[6]) if len(sys.argv) > 6 else (1 << 30) // parts
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 2  # 1 = successor records, 2 = parent forwarding
seqs = S7()
Gs, inbox, counts = [], [], []
for r in range(parts):
    G = m.PastarGPU(seqs)
    G.build_pair_tables()
    G.configure_hash(htype, shift)
    G.set_stream(torch.cuda.current_stream().cuda_stream)
    G.search_begin(parts, r, cap, batch, p2p=mode)
    Gs.append(G)
    inbox.append(torch.zeros(2 * parts * G.search_region_bytes(), dtype=torch.uint8, device="cuda"))
    counts.append(torch.zeros(2 * parts, dtype=torch.int64, device="cuda"))
for G in Gs:
    G.search_set_peers([t.data_ptr() for t in inbox])
    G.search_set_peer_counts([t.data_ptr() for t in counts], 2)
def run(n):
    for _ in range(n):
        for G in Gs:
            G.search_round_async()
        torch.cuda.synchronize()
        for G in Gs:
            G.search_insert_inbox_async()
        torch.cuda.synchronize()
run(rounds - 50)
for G in Gs:
    G.search_sync(); G.search_profile(True)
c0 = [G.search_status()[2] for G in Gs]
run(50)
for G in Gs:
    G.search_sync()
c1 = [G.search_status()[2] for G in Gs]
tot = {}
for r in range(parts):
    d = {k: (c1[r][k] - c0[r][k]) / 50 for k in ("select_ms", "claim_ms", "expand_ms", "insert_ms", "inbox_ms", "expansions", "generated", "pops", "survivors")}
    for k, v in d.items():
        tot.setdefault(k, []).append(v)
    if parts <= 4:
        print("part %d:" % r, {k: round(v, 4) if v < 10 else int(v) for k, v in d.items()}, flush=True)
step = [tot["select_ms"][r] + tot["claim_ms"][r] + tot["expand_ms"][r] + tot["insert_ms"][r] + tot["inbox_ms"][r] for r in range(parts)]
print("mode %d %s shift %d parts %d: kernel ms per round per partition max %.3f mean %.3f | expansions/round total %d | pops min %d | inbox max %.3f expand max %.3f | est. M exp/s at max %.0f" % (
    mode, htype, shift, parts, max(step), sum(step) / parts, sum(tot["expansions"]), min(tot["pops"]), max(tot["inbox_ms"]), max(tot["expand_ms"]), sum(tot["expansions"]) / max(step) / 1e3))
