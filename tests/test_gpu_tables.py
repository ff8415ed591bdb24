"""GPU parity (1): pairwise reverse-DP tables, calculate_h — through the C ABI, against the oracle."""
import numpy as np
import pytest

from conftest import CASES, S7, S8, random_seqs
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_tables_cell_for_cell(gpu_lib, name):
    seqs = CASES[name]
    with gpu_lib.PastarGPU(seqs) as g:
        g.build_pair_tables()
        k = 0
        for i in range(len(seqs) - 1):
            for j in range(i + 1, len(seqs)):
                ref = O.pair_table(seqs[i], seqs[j])
                got = g.pair_table(k)
                assert got.shape == ref.shape
                assert np.array_equal(got, ref), "pair %d (%d,%d) of %s" % (k, i, j, name)
                k += 1


@pytest.mark.parametrize("lens", [(1, 1, 1), (1, 40, 33), (32, 33, 31), (65, 64, 1), (100, 7, 300), (257, 1030, 2)])
def test_tables_ragged_lengths(gpu_lib, lens):
    seqs = [random_seqs(1, L, 100 + i)[0] for i, L in enumerate(lens)]
    with gpu_lib.PastarGPU(seqs, weights=None) as g:
        g.build_pair_tables()
        k = 0
        for i in range(len(seqs) - 1):
            for j in range(i + 1, len(seqs)):
                assert np.array_equal(g.pair_table(k), O.pair_table(seqs[i], seqs[j])), (lens, i, j)
                k += 1


def test_tables_affine_general(gpu_lib):
    """open != extension: the carried direction state must follow PairAlign.cpp:96-134 (oracle restates it)."""
    seqs = random_seqs(3, 90, 77)
    # the oracle's constants are the reference's 30/30; scale-check the general rule against a numpy DP instead
    with gpu_lib.PastarGPU(seqs, weights=None, gap_open=41, gap_ext=7) as g:
        g.build_pair_tables()
        cost = O.cost_table()
        k = 0
        for i in range(2):
            for j in range(i + 1, 3):
                s1, s2 = seqs[i], seqs[j]
                L1, L2 = len(s1), len(s2)
                M = np.zeros((L1 + 1, L2 + 1), dtype=np.int64)
                A = np.zeros((L1 + 1, L2 + 1), dtype=np.int64)  # 0 NoGap 1 GapX 2 GapY
                for b in range(L2 - 1, -1, -1):
                    M[L1, b] = 41 + (L2 - 1 - b) * 7
                    A[L1, b] = 2
                for a in range(L1 - 1, -1, -1):
                    M[a, L2] = 41 + (L1 - 1 - a) * 7
                    A[a, L2] = 1
                for a in range(L1 - 1, -1, -1):
                    for b in range(L2 - 1, -1, -1):
                        c0 = M[a + 1, b] + (7 if A[a + 1, b] == 1 else 41)
                        c1 = M[a, b + 1] + (7 if A[a, b + 1] == 2 else 41)
                        m, d = (c0, 1) if c0 < c1 else (c1, 2)
                        c2 = M[a + 1, b + 1] + cost[ord(s1[a]), ord(s2[b])]
                        if c2 < m:
                            m, d = c2, 0
                        M[a, b], A[a, b] = m, d
                assert np.array_equal(g.pair_table(k), M.astype(np.int32))
                k += 1


@pytest.mark.parametrize("make,cells", [(S7, 5271021), (S8, 28056028)])
def test_tables_full_size(gpu_lib, make, cells):
    """BASELINE sizes: every cell against the oracle (the C oracle does S8 in about half a second)."""
    seqs = make()
    with gpu_lib.PastarGPU(seqs, weights=None) as g:
        ms = g.build_pair_tables()
        assert ms > 0
        total = 0
        k = 0
        for i in range(len(seqs) - 1):
            for j in range(i + 1, len(seqs)):
                got = g.pair_table(k)
                total += got.size
                assert np.array_equal(got, O.pair_table(seqs[i], seqs[j]))
                k += 1
        assert total == cells


@pytest.mark.parametrize("name", ["test", "kinase", "fam6x80", "rnd7x120", "fam10x20", "fam16x5"])
def test_calculate_h(gpu_lib, name):
    seqs = CASES[name]
    P = O.Problem(seqs)
    rng = np.random.default_rng(5)
    lens = np.array([len(s) for s in seqs])
    coords = np.stack([rng.integers(0, lens + 1) for _ in range(500)]).astype(np.uint16)
    coords[0] = 0
    coords[1] = lens
    with gpu_lib.PastarGPU(seqs) as g:
        assert np.array_equal(g.w_int, P.int_weights())
        g.build_pair_tables()
        got = g.calculate_h(coords)
    ref = np.array([P.calculate_h(c) for c in coords], dtype=np.int32)
    assert np.array_equal(got, ref)


def test_state_errors(gpu_lib):
    g = gpu_lib.PastarGPU(CASES["PF08184"])
    with pytest.raises(gpu_lib.PastarError) as e:
        g.calculate_h(np.zeros((1, 3), dtype=np.uint16))
    assert e.value.code == 5  # PG_ERR_STATE: tables not built
    with pytest.raises(gpu_lib.PastarError) as e:
        g.configure_hash("FZORDER", 22)  # CoordHash.cpp:240-242 throws invalid_argument
    assert e.value.code == 6
    g.close()
    with pytest.raises(gpu_lib.PastarError):
        gpu_lib.PastarGPU(["AC", "AC"])  # N=2 is not in max_seq_helper.h


@pytest.mark.parametrize("lens", [(1100, 900, 40), (2200, 300, 5), (129, 128, 127), (16, 15, 17), (1, 700, 31)])
def test_tables_long_multi_pass(gpu_lib, lens):
    """More bands than warps (a warp takes several bands, rings wrap), 32-bit cells (30 * (L1 + L2) >= 65536) and band /
    chunk boundary lengths, through the linear-gap kernel."""
    seqs = [random_seqs(1, L, 300 + i)[0] for i, L in enumerate(lens)]
    with gpu_lib.PastarGPU(seqs, weights=None) as g:
        g.build_pair_tables()
        k = 0
        for i in range(len(seqs) - 1):
            for j in range(i + 1, len(seqs)):
                assert np.array_equal(g.pair_table(k), O.pair_table(seqs[i], seqs[j])), (lens, i, j)
                k += 1


def test_tables_wide_alphabet(gpu_lib):
    """Every residue code the 90 x 90 cost table admits (Cost.h:49), not just the 20 amino acids."""
    import random
    r = random.Random(9)
    seqs = ["".join(chr(r.randrange(33, 90)) for _ in range(L)) for L in (150, 140, 160)]
    with gpu_lib.PastarGPU(seqs, weights=None) as g:
        g.build_pair_tables()
        k = 0
        for i in range(2):
            for j in range(i + 1, 3):
                assert np.array_equal(g.pair_table(k), O.pair_table(seqs[i], seqs[j])), (i, j)
                k += 1


def _numpy_linear_dp(s1, s2, cost, gap):
    L1, L2 = len(s1), len(s2)
    M = np.zeros((L1 + 1, L2 + 1), dtype=np.int64)
    M[L1, :] = gap * (L2 - np.arange(L2 + 1))
    M[:, L2] = gap * (L1 - np.arange(L1 + 1))
    for a in range(L1 - 1, -1, -1):
        for b in range(L2 - 1, -1, -1):
            M[a, b] = min(M[a + 1, b] + gap, M[a, b + 1] + gap, M[a + 1, b + 1] + cost[ord(s1[a]), ord(s2[b])])
    return M.astype(np.int32)


@pytest.mark.parametrize("scale,general", [(1, True), (20, False), (3, False)])
def test_tables_custom_costs(gpu_lib, scale, general, monkeypatch):
    """A caller-supplied cost table: byte-sized costs take the linear-gap kernel, larger ones (x20: up to 500) and
    PG_DP_KERNEL=general take the general kernel; both must give the same cells as a plain DP."""
    if general:
        monkeypatch.setenv("PG_DP_KERNEL", "general")
    seqs = random_seqs(3, 70, 55)
    cost = (O.cost_table().astype(np.int32) * scale).astype(np.int32)
    with gpu_lib.PastarGPU(seqs, weights=None, cost=cost) as g:
        g.build_pair_tables()
        k = 0
        for i in range(2):
            for j in range(i + 1, 3):
                assert np.array_equal(g.pair_table(k), _numpy_linear_dp(seqs[i], seqs[j], cost, 30)), (scale, i, j)
                k += 1
