"""GPU parity: the Altschul pair weights with the pair loop (the reference's `primer`, pastar/WeightedSP.cpp:144-244) on
the device.  pg_host_weights is pinned float-bit-exact to the reference (tests/test_host.py, golden vectors);
pg_gpu_weights must give the same bit patterns."""
import numpy as np
import pytest

from conftest import CASES, S7, WIDE_CASES, random_seqs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES) + ["fam10x100"])
def test_gpu_weights_bit_exact(gpu_lib, name):
    seqs = CASES.get(name) or WIDE_CASES[name]
    host = gpu_lib.host_weights(seqs)
    dev = gpu_lib.gpu_weights(seqs)
    assert np.array_equal(host.view(np.uint32), dev.view(np.uint32)), (name, host, dev)


@pytest.mark.parametrize("n,L,seed", [(7, 500, 12345), (8, 998, 12345), (3, 1500, 4), (4, 1, 5), (3, 2, 6), (5, 33, 7)])
def test_gpu_weights_sizes(gpu_lib, n, L, seed):
    """BASELINE sizes (S7; S8 at the reference's own length limit 998 and beyond it), more rows than threads, tiny inputs."""
    seqs = random_seqs(n, L, seed)
    host = gpu_lib.host_weights(seqs)
    dev, ms = gpu_lib.gpu_weights(seqs, want_ms=True)
    assert ms > 0
    assert np.array_equal(host.view(np.uint32), dev.view(np.uint32))


def test_gpu_weights_ragged_and_context_default(gpu_lib):
    seqs = [random_seqs(1, L, 40 + i)[0] for i, L in enumerate((300, 17, 1100, 64))]
    assert np.array_equal(gpu_lib.host_weights(seqs).view(np.uint32), gpu_lib.gpu_weights(seqs).view(np.uint32))
    with gpu_lib.PastarGPU(CASES["kinase"]) as G:  # the context's default producer is the device path
        assert np.array_equal(G.weights_f.view(np.uint32), gpu_lib.host_weights(CASES["kinase"]).view(np.uint32))


def test_gpu_weights_too_long_is_refused(gpu_lib):
    with pytest.raises(gpu_lib.PastarError) as e:
        gpu_lib.gpu_weights(random_seqs(3, 7000, 1))
    assert e.value.code == 3  # PG_ERR_UNSUPPORTED: the caller takes pg_host_weights
