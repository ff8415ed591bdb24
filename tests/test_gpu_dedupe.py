"""GPU parity (3), unit level: the closed/open table (PAStar<N>::enqueue, pastar/PAStar.cpp:219-237; conditional_enqueue,
include/PriorityList.h:104-113; the pop-time closed check, PAStar.cpp:344-351) checked through pg_search_lookup, not
only through the final cost:
  * every node on the optimal path is in the table with exactly the prefix cost of the alignment and the move
    that leads to it (what the reference's ClosedList holds for the backtrace);
  * one node per round (batch 1) is plain A*: with a consistent heuristic nothing is ever reopened, every expansion
    closes one node, and the expansion count lies between the nodes with f < g* and those with f <= g*, as the serial
    oracle's does;
  * both key widths and both value widths."""
import numpy as np
import pytest

from conftest import CASES, KNOWN_OPT, WIDE_CASES, random_seqs
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def path_nodes(seqs, w_int, rows, cost):
    """(pos, g, move mask) of every node on the path an alignment describes (Node.cpp:129-152 cost model)."""
    n, cols = len(seqs), len(rows[0])
    pos, prev, g, out = [0] * n, [1] * n, 0, []
    for c in range(cols):
        mv = [0 if rows[i][c] == "-" else 1 for i in range(n)]
        for x in range(n - 1):
            for y in range(x + 1, n):
                if mv[x] and mv[y]:
                    t = int(cost[ord(rows[x][c]), ord(rows[y][c])])
                else:
                    t = 30  # GapOpen == GapExtension == GapGap (Cost.h:13)
                g += t * int(w_int[x][y])
        pos = [p + m for p, m in zip(pos, mv)]
        prev = mv
        out.append((list(pos), g, sum(m << i for i, m in enumerate(mv))))
    return out


@pytest.mark.parametrize("valw", ["auto", "8"])
@pytest.mark.parametrize("name,batch", [("PF08184", 64), ("test2", 16), ("fam5x60", 1024), ("fam7x30", 4096), ("fam8x20", 256), ("kinase", 16384),
                                        ("fam14x6", 256)])
def test_table_holds_the_optimal_path(gpu_lib, name, batch, valw, monkeypatch):
    if valw == "8":
        monkeypatch.setenv("PG_VALW", "8")
    seqs = CASES[name]
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        r = G.search(table_capacity=1 << 25, batch_target=batch)
        assert r["finished"] == 1
        nodes = path_nodes(seqs, G.w_int, r["rows"], gpu_lib.default_cost_table())
        assert nodes[-1][1] == r["g"]
        for pos, g, mask in nodes:
            hit = G.search_lookup(pos)
            assert hit is not None, (name, pos)
            assert hit[0] == g and hit[1] == mask, (name, pos, hit, g, mask)
        assert G.search_lookup([0] * len(seqs))[0] == 0  # the start node: g 0


def test_wide_key_table_holds_the_optimal_path(gpu_lib):
    seqs = WIDE_CASES["fam10x100"]  # 70 key bits: two-word block keys, 8-byte values
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        r = G.search(table_capacity=1 << 24, batch_target=4096)
        assert r["finished"] == 1 and r["g"] == KNOWN_OPT["fam10x100"]
        for pos, g, mask in path_nodes(seqs, G.w_int, r["rows"], gpu_lib.default_cost_table()):
            hit = G.search_lookup(pos)
            assert hit is not None and hit[0] == g and hit[1] == mask


@pytest.mark.parametrize("name", ["test", "test2", "PF08184", "fam5x60", "fam7x30", "rnd4x60"])
def test_batch_one_is_plain_astar(gpu_lib, name):
    seqs = CASES[name]
    ref = O.Problem(seqs).astar(want_rows=False)
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        a = G.search(table_capacity=1 << 22, batch_target=1)
        b = G.search(table_capacity=1 << 22, batch_target=1)
    assert a["finished"] == 1 and a["g"] == ref["g"]
    # deterministic: one pop per round, no races
    for k in ("expansions", "generated", "pops", "pushed", "inserted", "reopen", "closed_size", "open_size", "rounds"):
        assert a[k] == b[k], k
    assert a["reopen"] == 0 == ref.get("reopen", 0)          # consistent heuristic: a closed node is never improved
    assert a["pops"] <= a["rounds"]                           # one pop per round (rounds run in groups between host checks)
    assert a["closed_size"] in (a["expansions"], a["expansions"] + 1)  # every expansion closes one node (+ the goal)
    assert a["inserted"] == a["closed_size"] + a["open_size"]
    # any A* expands every node with f < g* and some with f == g*: the two drivers may differ only inside the last f layer
    assert abs(a["expansions"] - ref["expansions"]) <= max(8, ref["expansions"] // 2), (a["expansions"], ref["expansions"])


def test_s8_size_budgeted_search_wide_key(gpu_lib):
    """BASELINE configs[4] size (8 x 1000: 80 key bits) through the search kernels at a bench-like batch: the origin is
    closed with g 0, each of its 255 successors is in the table with a g no worse than the direct move's (getNeigh, pinned
    against the reference), at least one of them is closed, counters are consistent."""
    seqs = random_seqs(8, 1000, 12345)
    with gpu_lib.PastarGPU(seqs, weights=None) as G:
        G.build_pair_tables()
        r = G.search(table_capacity=1 << 26, batch_target=65536, max_expansions=400000)
        assert r["finished"] == 0 and r["expansions"] >= 400000
        assert r["inserted"] == r["open_size"] + r["closed_size"]
        hit = G.search_lookup([0] * 8)
        assert hit is not None and hit[0] == 0
        out, counts = G.expand_batch(G.make_nodes(np.zeros((1, 8), dtype=np.uint16), [0], [255]), 1)
        assert counts[0] == 255
        for rec in out[0, :255]:
            h2 = G.search_lookup([int(x) for x in rec["pos"][:8]])
            assert h2 is not None and h2[0] <= int(rec["g"]), (rec, h2)
