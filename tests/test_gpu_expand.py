"""GPU parity (2)+(3): batched getNeigh records (pos, f, g, parenti, owner) and get_id — against the oracle."""
import numpy as np
import pytest

from conftest import CASES, S7, random_parents, random_seqs
from oracle import oracle as O

pytestmark = pytest.mark.gpu

HASHES = [(1, "FZORDER", 12), (4, "FZORDER", 3), (3, "FZORDER", 0), (8, "PZORDER", 2), (5, "FSUM", 1), (7, "PSUM", 0),
          (6, "FZORDER", 5), (2, "FZORDER", 12), (64, "FZORDER", 21)]


def oracle_records(P, pos, g, par, vs, ht, sh):
    return [P.get_neigh(pos[k], g[k], par[k], vs, ht, sh) for k in range(len(pos))]


@pytest.mark.parametrize("name", list(CASES))
def test_expand_record_for_record(gpu_lib, name):
    seqs = CASES[name]
    n = len(seqs)
    K = 64 if n <= 10 else 6
    pos, g, par = random_parents(seqs, K, 7)
    P = O.Problem(seqs)
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        for (vs, ht, sh) in HASHES[:4] if n > 10 else HASHES:
            G.configure_hash(ht, sh)
            out, counts = G.expand_batch(G.make_nodes(pos, g, par), vs)
            ref = oracle_records(P, pos, g, par, vs, ht, sh)
            for k in range(K):
                r = ref[k]
                assert counts[k] == len(r), (name, k)
                got = out[k, :counts[k]]
                # ours: ascending mask; reference: bucket by owner, ascending mask inside (order-insensitive compare)
                r = r[np.argsort(r["parenti"], kind="stable")]
                for fld in ("pos", "f", "g", "parenti", "owner"):
                    assert np.array_equal(got[fld], r[fld]), (name, vs, ht, sh, k, fld)


def test_get_neigh_shim_order(gpu_lib):
    """The batch-of-one shim returns records in the reference's own order (Node.cpp:234-246)."""
    seqs = CASES["fam6x80"]
    P = O.Problem(seqs)
    pos, g, par = random_parents(seqs, 8, 3)
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        G.configure_hash("FZORDER", 2)
        for k in range(8):
            got = G.get_neigh(pos[k], g[k], par[k], 5)
            ref = P.get_neigh(pos[k], g[k], par[k], 5, "FZORDER", 2)
            for fld in ("pos", "f", "g", "parenti", "owner"):
                assert np.array_equal(got[fld], ref[fld])


def test_expand_full_size_s7(gpu_lib):
    """BASELINE size (7 x 500): a large batch, oracle-checked on a sample, structural checks on all of it."""
    seqs = S7()
    n, S = 7, 127
    K = 20000
    pos, g, par = random_parents(seqs, K, 11)
    P = O.Problem(seqs)
    with gpu_lib.PastarGPU(seqs) as G:
        G.build_pair_tables()
        out, counts = G.expand_batch(G.make_nodes(pos, g, par), 8)
        h_par = G.calculate_h(pos)
        # every successor: pos = parent + mask bits, f - g = h(pos), f >= f(parent) (consistent heuristic)
        interior = counts == S
        assert interior.sum() > K * 0.9
        sel = out[interior]
        masks = sel["parenti"]
        assert np.array_equal(masks, np.tile(np.arange(1, S + 1, dtype=np.int32), (sel.shape[0], 1)))
        bits = ((masks[..., None] >> np.arange(n)) & 1).astype(np.uint16)
        assert np.array_equal(sel["pos"], pos[interior][:, None, :] + bits)
        flat = sel.reshape(-1)
        hh = G.calculate_h(flat["pos"])
        assert np.array_equal(flat["f"] - flat["g"], hh)
        fpar = (g + h_par)[interior]
        assert (sel["f"] >= fpar[:, None]).all()
        assert np.array_equal(flat["owner"], G.owner(flat["pos"], 8))
        idx = np.random.default_rng(1).choice(K, 200, replace=False)
        for k in idx:
            r = P.get_neigh(pos[k], g[k], par[k], 8)
            r = r[np.argsort(r["parenti"], kind="stable")]
            got = out[k, :counts[k]]
            for fld in ("pos", "f", "g", "parenti", "owner"):
                assert np.array_equal(got[fld], r[fld])


def test_owner_all_hashes(gpu_lib):
    rng = np.random.default_rng(9)
    for n in (3, 4, 5, 6, 7, 8, 9, 10, 14, 16):
        seqs = random_seqs(n, 9, n)
        co = rng.integers(0, 65536, (400, n)).astype(np.uint16)
        co[:100] %= 1024
        with gpu_lib.PastarGPU(seqs, weights=None) as G:
            for ht in ("FZORDER", "PZORDER", "FSUM", "PSUM"):
                for sh in (0, 1, 5, 12, 13, 21):
                    G.configure_hash(ht, sh)
                    for size in (1, 2, 3, 5, 8, 13, 64):
                        got = G.owner(co, size)
                        ref = np.array([O.owner(c, ht, sh, size) for c in co], dtype=np.uint32)
                        assert np.array_equal(got, ref), (n, ht, sh, size)


def test_expand_empty_and_errors(gpu_lib):
    seqs = CASES["PF08184"]
    with gpu_lib.PastarGPU(seqs) as G:
        with pytest.raises(gpu_lib.PastarError):
            G.expand_batch(G.make_nodes(np.zeros((1, 3)), [0], [7]))  # tables not built
        G.build_pair_tables()
        out, counts = G.expand_batch(G.make_nodes(np.zeros((0, 3)), [], []))
        assert out.shape[0] == 0 and counts.shape[0] == 0
        # the final coordinate has no successors (borderCheck fails for every mask)
        out, counts = G.expand_batch(G.make_nodes([[59, 59, 59]], [5], [7]))
        assert counts[0] == 0
