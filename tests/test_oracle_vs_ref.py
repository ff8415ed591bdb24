"""CPU: the C oracle against the reference itself (oracle/_ref/pastar_ref = unmodified reference sources compiled in
place).  Skipped where the prebuilt binary is absent (it is built by __graft_entry__.build() when /root/reference exists)."""
import numpy as np
import pytest

from conftest import CASES, family_seqs, random_parents, random_seqs
from oracle import oracle as O
from oracle import refio as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/pastar_ref not built")

RAND = {"r5x50": random_seqs(5, 50, 41), "f4x90": family_seqs(4, 90, 42, 0.2, 0.05), "f16x5": CASES["fam16x5"],
        "f14x6": CASES["fam14x6"], "r3x998": random_seqs(3, 998, 43)}


@pytest.mark.parametrize("name", list(RAND))
def test_dump(name):
    seqs = RAND[name]
    d = R.dump(seqs)
    assert np.array_equal(O.cost_table(), d["cost"])
    assert np.array_equal(O.weights(seqs).view(np.uint32), d["weights"].view(np.uint32))
    k = 0
    for i in range(len(seqs) - 1):
        for j in range(i + 1, len(seqs)):
            assert np.array_equal(O.pair_table(seqs[i], seqs[j]), d["tables"][k])
            k += 1


@pytest.mark.parametrize("name", ["r5x50", "f4x90", "f14x6"])
def test_get_neigh(name):
    seqs = RAND[name]
    n = len(seqs)
    P = O.Problem(seqs)
    pos, g, par = random_parents(seqs, 6 if n <= 10 else 2, 5)
    for (vs, ht, sh) in [(1, "FZORDER", 12), (7, "FZORDER", 1), (4, "PZORDER", 3), (3, "FSUM", 2), (6, "PSUM", 0)]:
        ref = R.neigh(seqs, pos, g, par, vs, ht, sh)
        for k in range(len(pos)):
            mine = P.get_neigh(pos[k], g[k], par[k], vs, ht, sh)
            assert int(g[k]) + P.calculate_h(pos[k]) == ref[k][0]
            assert np.array_equal(mine, ref[k][1])


def test_astar_and_partitioned_driver():
    seqs = RAND["f4x90"]
    P = O.Problem(seqs)
    a = P.astar(want_rows=False)
    r = R.astar(seqs)
    assert a["g"] == r["g"] and a["expansions"] == r["expansions"] and a["generated"] == r["generated"]
    p = R.pastar(seqs, 4)  # T-thread hash-partitioned driver (PAStar.cpp:319-547 restated): same optimal cost
    assert p["finished"] == 1 and p["g"] == r["g"]
