"""CPU, world_size 2 (and 4), gloo: the hash-partitioned driver (exchange by owner, allreduce(min) stop, distributed backtrace)
with a TEST-ONLY engine that stands in for the CUDA kernels (oracle getNeigh + a dict as closed/open table).  The product
engine is CudaEngine; this covers the host-side logic of the N>1 path without a GPU."""
import heapq
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CASES, KNOWN_OPT, weighted_sp_score

INT_MAX = 2**31 - 1


class OracleEngine:
    """Same interface as mpi_pastar_msa_b200.dist.CudaEngine, CPU only, for tests."""

    def __init__(self, seqs, n_parts, part, batch, hash_type="FZORDER", shift=3):
        from oracle import oracle as O
        self.O, self.P = O, O.Problem(seqs)
        self.n, self.n_parts, self.part, self.batch = len(seqs), n_parts, part, batch
        self.ht, self.sh = hash_type, shift
        self.final = tuple(len(s) for s in seqs)
        self.rec = np.dtype([("pos", np.uint16, (self.n,)), ("g", np.int32), ("f", np.int32), ("parenti", np.int32)])
        self.table = {}   # pos -> [g, parenti, open]
        self.heap = []    # (f, g, pos)
        self.best_goal = INT_MAX
        self.cnt = {"expansions": 0, "generated": 0, "pops": 0}
        self.device = torch.device("cpu")
        if part == 0:
            start = tuple([0] * self.n)
            self._offer(start, 0, self.P.calculate_h(start), (1 << self.n) - 1)

    def _offer(self, pos, g, f, par):
        e = self.table.get(pos)
        if e is not None and g >= e[0]:
            return
        self.table[pos] = [g, par, True]
        if pos == self.final:
            self.best_goal = min(self.best_goal, g)
        heapq.heappush(self.heap, (f, g, pos))

    def round(self, f_limit):
        out = [[] for _ in range(self.n_parts)]
        lim = min(f_limit, self.best_goal)
        popped = 0
        while self.heap and popped < self.batch and self.heap[0][0] < lim:
            f, g, pos = heapq.heappop(self.heap)
            self.cnt["pops"] += 1
            e = self.table[pos]
            if g != e[0] or not e[2]:
                continue
            e[2] = False
            popped += 1
            if pos == self.final:
                continue
            self.cnt["expansions"] += 1
            for s in self.P.get_neigh(pos, g, e[1], self.n_parts, self.ht, self.sh):
                self.cnt["generated"] += 1
                sp = tuple(int(x) for x in s["pos"])
                if sp == self.final:
                    self.best_goal = min(self.best_goal, int(s["g"]))
                if int(s["owner"]) == self.part:
                    self._offer(sp, int(s["g"]), int(s["f"]), int(s["parenti"]))
                else:
                    out[int(s["owner"])].append((s["pos"], int(s["g"]), int(s["f"]), int(s["parenti"])))
        bufs = []
        for lst in out:
            a = np.zeros(len(lst), dtype=self.rec)
            for i, (p, g, f, par) in enumerate(lst):
                a[i] = (p, g, f, par)
            bufs.append(torch.from_numpy(a.view(np.uint8).reshape(-1).copy()))
        return bufs

    def insert(self, buf):
        a = buf.numpy().view(self.rec)
        for r in a:
            self._offer(tuple(int(x) for x in r["pos"]), int(r["g"]), int(r["f"]), int(r["parenti"]))

    def status(self):
        while self.heap:  # drop stale heads so min_open_f is exact
            f, g, pos = self.heap[0]
            e = self.table[pos]
            if g == e[0] and e[2]:
                break
            heapq.heappop(self.heap)
        return (self.heap[0][0] if self.heap else INT_MAX), self.best_goal, dict(self.cnt)

    def lookup(self, pos):
        e = self.table.get(tuple(int(x) for x in pos))
        return (e[0], e[1]) if e else None

    def empty(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8)


def _worker(rank, world, port, name, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpi_pastar_msa_b200.dist import PartitionedSearch
    from oracle import oracle as O
    seqs = CASES[name]
    eng = OracleEngine(seqs, world, rank, batch)
    drv = PartitionedSearch(eng, dist, seqs, lambda pos: O.owner(pos, eng.ht, eng.sh, world))
    res = drv.run()
    res["bytes_sent"] = drv.bytes_sent
    res["table"] = len(eng.table)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,batch", [("PF08184", 4), ("test2", 16), ("fam5x60", 64), ("fam3x300", 256)])
def test_partitioned_search_two_ranks(name, batch):
    from oracle import oracle as O
    seqs = CASES[name]
    ref = O.Problem(seqs).astar(want_rows=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        res = got[r]
        assert res["finished"] == 1 and res["g"] == ref["g"], (name, r, res["g"], ref["g"])
        if name in KNOWN_OPT:
            assert res["g"] == KNOWN_OPT[name]
        assert weighted_sp_score(seqs, O.Problem(seqs).int_weights(), res["rows"]) == ref["g"]
    assert got[0]["rows"] == got[1]["rows"]
    assert got[0]["expansions"] == got[1]["expansions"] >= 1   # allreduced totals agree on every rank
    assert got[0]["bytes_sent"] + got[1]["bytes_sent"] > 0      # successors really crossed partitions
    assert got[0]["table"] > 0 and got[1]["table"] > 0          # both partitions own part of the state space


@pytest.mark.parametrize("name,batch", [("fam5x60", 16), ("PF08184", 2)])
def test_partitioned_search_four_ranks(name, batch):
    """G = 4: two owner bits (FZORDER: `(Z >> shift) & 3`), three peers per rank in the all-to-all, stop test over four partitions."""
    from oracle import oracle as O
    seqs = CASES[name]
    ref = O.Problem(seqs).astar(want_rows=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 4
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        res = got[r]
        assert res["finished"] == 1 and res["g"] == ref["g"], (name, r, res["g"], ref["g"])
        assert res["rows"] == got[0]["rows"] and res["expansions"] == got[0]["expansions"]
    assert weighted_sp_score(seqs, O.Problem(seqs).int_weights(), got[0]["rows"]) == ref["g"]
    assert sum(got[r]["bytes_sent"] for r in range(world)) > 0
    assert sum(1 for r in range(world) if got[r]["table"] > 0) >= 2   # the state space is really split


class ChainedOracleEngine(OracleEngine):
    """Stand-in for CudaEngineP2P: the engine itself moves the records (here: a gloo all-to-all) and the driver chains
    several rounds between status exchanges, passing a best-goal bound that is a few rounds old."""

    async_rounds = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.bytes_sent = 0

    def round_and_exchange(self, f_limit, d):
        out = self.round(f_limit)
        send_n = torch.tensor([b.numel() for b in out], dtype=torch.int64)
        recv_n = torch.empty_like(send_n)
        d.all_to_all_single(recv_n, send_n)
        inbox = torch.empty(int(recv_n.sum()), dtype=torch.uint8)
        d.all_to_all_single(inbox, torch.cat(out) if int(send_n.sum()) else torch.empty(0, dtype=torch.uint8),
                            output_split_sizes=recv_n.tolist(), input_split_sizes=send_n.tolist())
        self.bytes_sent += int(send_n.sum())
        self.insert(inbox)


def _worker_chained(rank, world, port, name, batch, per_status, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpi_pastar_msa_b200.dist import PartitionedSearch
    from oracle import oracle as O
    seqs = CASES[name]
    eng = ChainedOracleEngine(seqs, world, rank, batch)
    drv = PartitionedSearch(eng, dist, seqs, lambda pos: O.owner(pos, eng.ht, eng.sh, world))
    drv.rounds_per_status = per_status
    res = drv.run()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,batch,per_status", [("PF08184", 4, 4), ("fam5x60", 32, 3), ("test2", 8, 8)])
def test_chained_rounds_keep_the_optimum(name, batch, per_status):
    """Device-driven engines run `rounds_per_status` rounds between stop tests with a stale best-goal bound: the rounds
    past the optimum only pop nodes with f >= g*, so the result must not change."""
    from oracle import oracle as O
    seqs = CASES[name]
    ref = O.Problem(seqs).astar(want_rows=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_chained, args=(r, 2, port, name, batch, per_status, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r]["finished"] == 1 and got[r]["g"] == ref["g"], (name, r, got[r]["g"], ref["g"])
        assert got[r]["rounds"] % per_status == 0
        assert weighted_sp_score(seqs, O.Problem(seqs).int_weights(), got[r]["rows"]) == ref["g"]
