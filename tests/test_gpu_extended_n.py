"""sequence counts outside the reference's template set {3..10, 14, 16} (pastar/include/max_seq_helper.h:9-19)
behind an explicit opt-in.  Same parity bar as everything else: tables, weights, successor records and the optimal cost
against the oracle (whose C restatement is generic in N)."""
import numpy as np
import pytest

from conftest import family_seqs, random_parents, weighted_sp_score
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def extended(gpu_lib):
    gpu_lib.allow_extended_n(True)
    yield gpu_lib
    gpu_lib.allow_extended_n(False)


def test_rejected_without_the_opt_in(gpu_lib):
    gpu_lib.allow_extended_n(False)
    with pytest.raises(gpu_lib.PastarError) as e:
        gpu_lib.PastarGPU(family_seqs(11, 6, 1), weights=None)
    assert e.value.code == 1  # PG_ERR_ARG, as for any N the reference cannot run


@pytest.mark.parametrize("n,L,seed", [(11, 8, 61), (12, 7, 62), (13, 6, 63), (15, 5, 64)])
def test_extended_n_parity(extended, n, L, seed):
    m = extended
    seqs = family_seqs(n, L, seed, 0.2, 0.0)
    P = O.Problem(seqs)
    with m.PastarGPU(seqs) as G:
        assert np.array_equal(G.w_int, P.int_weights())
        G.build_pair_tables()
        for k in range(G.npairs):
            assert np.array_equal(G.pair_table(k), P.table(k))
        pos, g, par = random_parents(seqs, 6, seed)
        out, counts = G.expand_batch(G.make_nodes(pos, g, par), 5)
        for k in range(len(pos)):
            ref = P.get_neigh(pos[k], int(g[k]), int(par[k]), 5)
            ref = ref[np.argsort(ref["parenti"], kind="stable")]
            got = out[k, :counts[k]]
            for fld in ("pos", "f", "g", "parenti", "owner"):
                assert np.array_equal(got[fld], ref[fld]), (n, k, fld)
        ref = P.astar(want_rows=False)
        r = G.search(table_capacity=1 << 22, batch_target=256)
        assert r["finished"] == 1 and r["g"] == ref["g"], (n, r, ref)
        assert weighted_sp_score(seqs, G.w_int, r["rows"]) == ref["g"]
        with pytest.raises(m.PastarError) as e:  # the partitioned kernels are not built for these N
            G.search_begin(2, 0, 1 << 20, 256, p2p=2)
        assert e.value.code == 3
