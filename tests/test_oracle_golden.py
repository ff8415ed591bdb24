"""CPU: the C oracle (oracle/pastar_oracle.c) against the golden vectors generated from the unmodified reference."""
import os

import numpy as np
import pytest

from conftest import CASES, GOLDEN, KNOWN_OPT, S7
from oracle import oracle as O

G = np.load(os.path.join(GOLDEN, "reference_golden.npz"))
ALL = dict(CASES)
ALL["S7"] = S7()
HASH_CFGS = [(1, "FZORDER", 12), (5, "FZORDER", 3)]


def test_cost_table_matches_reference():
    ct = O.cost_table()
    assert np.array_equal(ct, G["cost_table"])
    # spot values from Cost.cpp: C/C 5, W/W 0, DASH/A 12, DASH/C unset, B unset
    assert ct[ord("C"), ord("C")] == 5 and ct[ord("W"), ord("W")] == 0 and ct[ord("-"), ord("A")] == 12
    assert ct[ord("-"), ord("C")] == 0 and ct[ord("A"), ord("B")] == 0 and ct[0, ord("A")] == 0


@pytest.mark.parametrize("name", list(ALL))
def test_sequences_are_the_generated_ones(name):
    assert list(G[name + "/seqs"]) == ALL[name]


@pytest.mark.parametrize("name", list(ALL))
def test_weights_bit_exact(name):
    w = O.weights(ALL[name])
    assert np.array_equal(w.view(np.uint32), G[name + "/weights_f32"].view(np.uint32))


@pytest.mark.parametrize("name", list(ALL))
def test_tables(name):
    seqs = ALL[name]
    k = 0
    for i in range(len(seqs) - 1):
        for j in range(i + 1, len(seqs)):
            t = O.pair_table(seqs[i], seqs[j])
            assert int(t.astype(np.int64).sum()) == int(G[name + "/table_sums"][k])
            assert [t[0, 0], t[0, -1], t[-1, 0], t[t.shape[0] // 2, t.shape[1] // 3]] == list(G[name + "/table_corner"][k])
            k += 1


def test_survey_known_answers():
    """Pins recorded in SURVEY §4 from the reference's own arithmetic."""
    P = O.Problem(CASES["PF08184"])
    assert [int(P.table(k)[0, 0]) for k in range(3)] == [654, 666, 666]
    assert list(P.int_weights()[np.triu_indices(3, 1)]) == [16, 13, 8]
    s = P.get_neigh([0, 0, 0], 0, 7)
    by = {int(r["parenti"]): (int(r["g"]), int(r["f"] - r["g"])) for r in s}
    assert by[1] == (1110, 24943) and by[3] == (838, 24599) and by[6] == (974, 24839) and by[7] == (481, 23969)
    K = O.Problem(CASES["kinase"])
    assert K.calculate_h([0] * 5) == 408457
    assert [int(K.table(k)[0, 0]) for k in range(10)] == [4639, 4564, 4694, 4663, 4789, 4763, 4754, 4665, 4674, 4755]
    assert [int(K.table(k).astype(np.int64).sum()) for k in range(10)] == [304148314, 283444511, 299286238, 298166982, 298837845,
                                                                          315367291, 312951372, 293817043, 291379635, 308786947]
    by = {int(r["parenti"]): (int(r["g"]), int(r["f"] - r["g"])) for r in K.get_neigh([0] * 5, 0, 31)}
    assert by[1] == (2610, 408925) and by[3] == (2475, 408692) and by[30] == (1915, 408018) and by[31] == (1402, 407055)
    T = O.Problem(CASES["test"])
    assert list(T.int_weights()[0, 1:]) == [341, 187, 231, 113, 91, 148, 91]


@pytest.mark.parametrize("name", [n for n in ALL if len(ALL[n]) <= 10])
def test_get_neigh_records(name):
    seqs = ALL[name]
    P = O.Problem(seqs)
    pos, g, par = G[name + "/parents_pos"], G[name + "/parents_g"], G[name + "/parents_par"]
    for ci, (vs, ht, sh) in enumerate(HASH_CFGS):
        counts = G[name + "/neigh%d_count" % ci]
        off = 0
        for k in range(len(pos)):
            s = P.get_neigh(pos[k], g[k], par[k], vs, ht, sh)
            assert len(s) == counts[k]
            assert int(g[k]) + P.calculate_h(pos[k]) == int(G[name + "/neigh%d_fpar" % ci][k])
            for fld in ("pos", "f", "g", "parenti", "owner"):
                assert np.array_equal(s[fld], G[name + "/neigh%d_%s" % (ci, fld)][off:off + counts[k]]), (name, k, fld)
            off += counts[k]


def test_owner_hashes():
    for n in (3, 5, 7, 8, 10, 16):
        co = G["owner/%d/coords" % n]
        for ht in ("FZORDER", "PZORDER", "FSUM", "PSUM"):
            for sh in (0, 5, 12, 21):
                for size in (1, 3, 8, 64):
                    ref = G["owner/%d/%s/%d/%d" % (n, ht, sh, size)]
                    got = np.array([O.owner(c, ht, sh, size) for c in co], dtype=np.uint32)
                    assert np.array_equal(got, ref), (n, ht, sh, size)
    assert O.owner([1, 2, 3], "FZORDER", 22, 4) == 0xFFFFFFFF  # CoordHash.cpp:240-242 throws


@pytest.mark.parametrize("name", ["test", "test2", "PF08184", "rnd4x60", "fam6x80", "fam3x300", "fam5x60", "fam4x150", "fam7x30",
                                  "fam8x20"])
def test_astar_optimal_cost(name):
    from conftest import weighted_sp_score
    seqs = ALL[name]
    P = O.Problem(seqs)
    r = P.astar()
    ref = G[name + "/astar"]
    assert r["finished"] == 1 and r["g"] == int(ref[0])
    assert r["expansions"] == int(ref[1]) and r["generated"] == int(ref[2])  # same tie-breaking as the restated serial driver
    if name in KNOWN_OPT:
        assert r["g"] == KNOWN_OPT[name]
    assert weighted_sp_score(seqs, P.int_weights(), r["rows"]) == r["g"]


def test_astar_budget():
    r = O.Problem(CASES["kinase"]).astar(budget=2000, want_rows=False)
    assert r["finished"] == 0 and r["pops"] == 2000
