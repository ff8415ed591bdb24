"""GPU parity (4): the hash-partitioned search with G logical partitions on ONE GPU (one context per partition, the
all-to-all replaced by handing each outbox's device pointer to its owner's insert) — same optimal cost as the oracle.
This is the N>1 data path (owner routing in the expand kernel, outbox records, insert kernel, distributed stop and
backtrace) without needing G devices; NCCL itself is exercised by tools/multi_gpu_check.py under torchrun."""
import numpy as np
import pytest

from conftest import CASES, KNOWN_OPT, weighted_sp_score
from oracle import oracle as O

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


def partitioned_search(m, seqs, parts, batch, hash_type, shift, cap=1 << 22):
    Gs = []
    for r in range(parts):
        G = m.PastarGPU(seqs)
        G.build_pair_tables()
        G.configure_hash(hash_type, shift)
        G.search_begin(parts, r, cap, batch)
        Gs.append(G)
    best, rounds, sent = INT_MAX, 0, 0
    while True:
        for G in Gs:
            G.search_round(best)
        boxes = [[G.search_outbox(d) if d != r else (0, 0) for d in range(parts)] for r, G in enumerate(Gs)]
        for r in range(parts):
            for d in range(parts):
                ptr, n = boxes[r][d]
                if n:
                    Gs[d].search_insert_dev(ptr, n)
                    sent += n
        st = [G.search_status() for G in Gs]
        mn = min(s[0] for s in st)
        best = min(s[1] for s in st)
        rounds += 1
        assert rounds < 200000
        if mn >= best or mn == INT_MAX:
            break
    res = {"g": best if best != INT_MAX else -1, "rounds": rounds, "sent": sent,
           "expansions": sum(s[2]["expansions"] for s in st)}
    # distributed backtrace: the owner of each coordinate answers
    if best != INT_MAX:
        n = len(seqs)
        pos = [len(s) for s in seqs]
        cols = []
        while any(pos):
            own = int(Gs[0].owner(np.array(pos, dtype=np.uint16), parts)[0])
            hit = Gs[own].search_lookup(pos)
            assert hit is not None, pos
            cols.append(hit[1])
            pos = [p - ((hit[1] >> i) & 1) for i, p in enumerate(pos)]
        rows, at = [[] for _ in range(n)], [0] * n
        for mask in reversed(cols):
            for i in range(n):
                if (mask >> i) & 1:
                    rows[i].append(seqs[i][at[i]])
                    at[i] += 1
                else:
                    rows[i].append("-")
        res["rows"] = ["".join(r) for r in rows]
    for G in Gs:
        G.search_end()
        G.close()
    return res


@pytest.mark.parametrize("name,parts,batch,ht,sh", [
    ("PF08184", 2, 64, "FZORDER", 3), ("test2", 2, 256, "FSUM", 1), ("fam5x60", 2, 1024, "FZORDER", 2),
    ("fam5x60", 3, 64, "FZORDER", 0), ("fam8x20", 4, 4096, "PZORDER", 1), ("fam4x150", 8, 4096, "FZORDER", 12),
    ("fam6x80", 5, 512, "PSUM", 2), ("test", 4, 16, "FZORDER", 1)])
def test_partitioned_optimal_cost(gpu_lib, name, parts, batch, ht, sh):
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    r = partitioned_search(gpu_lib, seqs, parts, batch, ht, sh)
    assert r["g"] == ref, (name, parts, r)
    w = gpu_lib.host_weights(seqs).astype(np.int32)
    assert weighted_sp_score(seqs, w, r["rows"]) == ref
    if parts > 1 and name != "test":
        assert r["sent"] > 0  # successors really crossed partitions


def test_partitioned_kinase_two_parts(gpu_lib):
    seqs = CASES["kinase"]
    r = partitioned_search(gpu_lib, seqs, 2, 32768, "FZORDER", 12, cap=1 << 26)
    assert r["g"] == 421546


def p2p_search(m, seqs, parts, batch, hash_type, shift, mode, cap=1 << 22, device_sync=False):
    """The device-driven P2P rounds (pg_search_round_async / pg_search_insert_inbox_async) with G logical partitions on
    ONE GPU: every partition's inbox and count array are plain device buffers of this process, so "peer-mapped" is
    simply their address, and the cross-GPU barrier is a device synchronise.  mode 1 = successor records stored by the
    expand kernel, mode 2 = parent forwarding (forward_kernel + owner-filtered expansion of the received parents)."""
    import torch
    Gs, inbox, counts = [], [], []
    for r in range(parts):
        G = m.PastarGPU(seqs)
        G.build_pair_tables()
        G.configure_hash(hash_type, shift)
        G.set_stream(torch.cuda.current_stream().cuda_stream)
        G.search_begin(parts, r, cap, batch, p2p=mode)
        Gs.append(G)
        inbox.append(torch.zeros(2 * parts * G.search_region_bytes(), dtype=torch.uint8, device="cuda"))
        counts.append(torch.zeros(2 * parts, dtype=torch.int64, device="cuda"))
    for G in Gs:
        G.search_set_peers([t.data_ptr() for t in inbox])
        G.search_set_peer_counts([t.data_ptr() for t in counts], 2)
        if device_sync:  # stamped counts + device-side wait instead of a barrier between the two halves of a round
            G.search_set_device_sync(True)
    best, rounds = INT_MAX, 0
    while True:
        for _ in range(3):  # a few rounds between status checks, as the torchrun driver does
            for G in Gs:
                G.search_round_async(best)
            torch.cuda.synchronize()
            for G in Gs:
                G.search_insert_inbox_async()
            torch.cuda.synchronize()
            rounds += 1
        for G in Gs:
            G.search_sync()
        st = [G.search_status() for G in Gs]
        mn = min(s[0] for s in st)
        best = min(s[1] for s in st)
        assert rounds < 200000
        if mn >= best or mn == INT_MAX:
            break
    res = {"g": best if best != INT_MAX else -1, "rounds": rounds, "expansions": sum(s[2]["expansions"] for s in st),
           "generated": sum(s[2]["generated"] for s in st)}
    for G in Gs:
        G.search_end()
        G.close()
    return res


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name,parts,batch,ht,sh", [
    ("PF08184", 2, 64, "FZORDER", 3), ("test2", 2, 256, "FSUM", 1), ("fam5x60", 3, 64, "FZORDER", 0),
    ("fam8x20", 4, 4096, "PZORDER", 1), ("fam4x150", 8, 4096, "FZORDER", 12), ("fam6x80", 5, 512, "PSUM", 2),
    ("test", 4, 16, "FZORDER", 1), ("fam5x60", 16, 1024, "FZORDER", 2)])
def test_p2p_modes_optimal_cost(gpu_lib, name, parts, batch, ht, sh, mode):
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    r = p2p_search(gpu_lib, seqs, parts, batch, ht, sh, mode)
    assert r["g"] == ref, (name, parts, mode, r)


def test_forwarding_generates_each_successor_once(gpu_lib):
    """Parent forwarding must neither lose nor duplicate successors: every valid successor of an expansion is counted
    exactly once, by its owner, so successors per expansion must match the successor-record mode (the expansion counts
    themselves may differ by a few reopened nodes: the insertion order inside a round is not deterministic)."""
    seqs = CASES["fam5x60"]
    r1 = p2p_search(gpu_lib, seqs, 4, 1024, "FZORDER", 2, 1)
    r2 = p2p_search(gpu_lib, seqs, 4, 1024, "FZORDER", 2, 2)
    assert r1["g"] == r2["g"]
    assert abs(r1["expansions"] - r2["expansions"]) <= 0.05 * r1["expansions"]
    per1, per2 = r1["generated"] / r1["expansions"], r2["generated"] / r2["expansions"]
    assert abs(per1 - per2) <= 0.01 * per1, (r1, r2)


@pytest.mark.parametrize("name,parts,batch,ht,sh", [
    ("PF08184", 2, 64, "FZORDER", 3), ("fam5x60", 3, 256, "FZORDER", 0), ("fam8x20", 4, 4096, "PZORDER", 1),
    ("fam4x150", 8, 4096, "FZORDER", 12), ("fam6x80", 5, 512, "FSUM", 2)])
def test_multi_search_one_process(gpu_lib, name, parts, batch, ht, sh):
    """pg_multi_search (what `pastar -g G` calls): G partitions driven by one process; on this one-GPU box the G contexts
    share the device, the exchange code path (forwarded parents, event barrier, owner lookups for the backtrace) is the
    same.  Optimal cost, a valid alignment, and per-partition counters that add up."""
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    Gs = []
    for _ in range(parts):
        G = gpu_lib.PastarGPU(seqs)
        G.build_pair_tables()
        G.configure_hash(ht, sh)
        Gs.append(G)
    tot, per = gpu_lib.multi_search(Gs, table_capacity=1 << 22, batch_target=batch)
    assert tot["finished"] == 1 and tot["g"] == ref, tot
    w = gpu_lib.host_weights(seqs).astype(np.int32)
    assert weighted_sp_score(seqs, w, tot["rows"]) == ref
    assert sum(p["expansions"] for p in per) == tot["expansions"] and sum(p["pops"] for p in per) == tot["pops"]
    assert tot["closed_size"] == sum(p["closed_size"] for p in per) > 0
    for G in Gs:
        G.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_p2p_modes_wide_key(gpu_lib, mode):
    """The two-word key (KEYW = 2) through successor records (32-byte pg_xrec) and parent forwarding (24-byte parents)."""
    from conftest import WIDE_CASES
    r = p2p_search(gpu_lib, WIDE_CASES["fam10x100"], 4, 4096, "FZORDER", 6, mode, cap=1 << 24)
    assert r["g"] == KNOWN_OPT["fam10x100"], r


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("name,parts,batch,ht,sh", [
    ("PF08184", 2, 64, "FZORDER", 12), ("kinase", 2, 16384, "FZORDER", 12), ("fam5x60", 2, 256, "FZORDER", 0),
    ("fam14x6", 2, 256, "FZORDER", 1), ("fam16x5", 2, 128, "FZORDER", 0), ("fam8x20", 4, 4096, "PZORDER", 1),
    ("fam4x150", 8, 4096, "FZORDER", 12)])
def test_multi_search_distinct_devices(gpu_lib, name, parts, batch, ht, sh):
    """pg_multi_search with one context per REAL device (needs >= parts GPUs: `gpurun --gpus N`): peer-mapped inboxes over
    NVLink, cross-device visibility of forward_kernel's stores, the stream-wait barrier.  N = 14 / 16 need more than
    48 KB of dynamic shared memory in the expand kernel: the opt-in is per device and must be made on every device."""
    if _n_devices() < parts:
        pytest.skip("needs %d GPUs" % parts)
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    Gs = []
    for dev in range(parts):
        G = gpu_lib.PastarGPU(seqs, device=dev)
        G.build_pair_tables()
        G.configure_hash(ht, sh)
        Gs.append(G)
    tot, per = gpu_lib.multi_search(Gs, table_capacity=1 << 24, batch_target=batch)
    assert tot["finished"] == 1 and tot["g"] == ref, tot
    w = gpu_lib.host_weights(seqs).astype(np.int32)
    assert weighted_sp_score(seqs, w, tot["rows"]) == ref
    assert gpu_lib.rescore_alignment(seqs, w, tot["rows"]) == ref
    assert sum(p["expansions"] for p in per) == tot["expansions"]
    for G in Gs:
        G.close()


@pytest.mark.parametrize("name", ["fam14x6", "fam16x5"])
def test_multi_search_wide_n_one_device(gpu_lib, name):
    """N = 14 / 16 through pg_multi_search on one device (two contexts): each context opts its kernels in to > 48 KB of
    dynamic shared memory itself (the cache is per context, not per process)."""
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    Gs = []
    for _ in range(2):
        G = gpu_lib.PastarGPU(seqs)
        G.build_pair_tables()
        G.configure_hash("FZORDER", 1)
        Gs.append(G)
    tot, _ = gpu_lib.multi_search(Gs, table_capacity=1 << 22, batch_target=256)
    assert tot["finished"] == 1 and tot["g"] == ref, tot
    for G in Gs:
        G.close()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name,parts,batch,ht,sh", [("PF08184", 2, 64, "FZORDER", 3), ("fam5x60", 3, 64, "FZORDER", 0), ("fam8x20", 4, 4096, "PZORDER", 1),
                                                    ("fam4x150", 8, 4096, "FZORDER", 12), ("kinase", 2, 16384, "PZORDER", 6)])
def test_p2p_modes_device_sync(gpu_lib, name, parts, batch, ht, sh, mode):
    """Data-flow synchronisation (pg_search_set_device_sync): counts carry the exchange round and the receiving half of a
    round waits for them on the device.  Same optimal costs; on one device every publish precedes every wait in stream
    order, the cross-device ordering is exercised by bench.py --gpus N's parity gate."""
    seqs = CASES[name]
    ref = KNOWN_OPT.get(name) or O.Problem(seqs).astar(want_rows=False)["g"]
    r = p2p_search(gpu_lib, seqs, parts, batch, ht, sh, mode, cap=1 << 24, device_sync=True)
    assert r["g"] == ref, (name, parts, mode, r)
