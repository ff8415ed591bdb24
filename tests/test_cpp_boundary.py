"""The drop-in boundary bound from C++: tests/cpp/boundary_test.cpp includes the host mirror (pastar_host.hpp) and calls
Node<N>::getNeigh, Coord<N>::get_id and HeuristicHPair::calculate_h exactly as the reference's callers do
(pastar/include/Node.h:37, Coord.h:49, HeuristicHPair.h:18-23); its output is compared with the oracle (pinned to the
reference compiled in place, tests/test_oracle_vs_ref.py) record for record."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES, ROOT, random_parents
from oracle import oracle as O

BIN = os.path.join(ROOT, "mpi_pastar_msa_b200", "bin", "boundary_test")


def test_boundary_program_is_built():
    assert os.path.exists(BIN), "python -m mpi_pastar_msa_b200.build"


@pytest.mark.gpu
@pytest.mark.parametrize("name,vec,ht,sh", [("PF08184", 4, "FZORDER", 3), ("kinase", 8, "FZORDER", 12), ("fam7x30", 5, "PZORDER", 2),
                                            ("fam8x20", 3, "FSUM", 1), ("test", 8, "PSUM", 0), ("fam14x6", 2, "FZORDER", 1)])
def test_cpp_getneigh_getid_calculate_h(tmp_path, name, vec, ht, sh):
    seqs = CASES[name]
    n = len(seqs)
    fa = tmp_path / "in.fasta"
    fa.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    pos, g, par = random_parents(seqs, 12, 7)
    stdin = "".join(" ".join(str(int(x)) for x in pos[k]) + " %d %d\n" % (int(g[k]), int(par[k])) for k in range(len(pos)))
    r = subprocess.run([BIN, str(fa), str(vec), ht, str(sh)], input=stdin.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    P = O.Problem(seqs)
    lines = [l for l in r.stdout.decode().splitlines() if l and l[0] in "HISE" and not l.startswith("Starting")]
    it = iter(lines)
    for k in range(len(pos)):
        h = next(it).split()
        assert h[0] == "H" and int(h[1]) == P.calculate_h(pos[k])
        i = next(it).split()
        assert i[0] == "ID" and int(i[1]) == int(O.owner(pos[k], ht, sh, vec))
        got = []
        for l in it:
            if l == "E":
                break
            t = l.split()
            assert t[0] == "S"
            got.append(tuple(int(x) for x in t[1:]))
        ref = P.get_neigh(pos[k], int(g[k]), int(par[k]), vec, ht, sh)
        want = [tuple([int(x["owner"])] + [int(c) for c in x["pos"][:n]] + [int(x["f"]), int(x["g"]), int(x["parenti"])]) for x in ref]
        # the reference appends bucket by bucket (owner), ascending move mask inside a bucket
        assert got == sorted(want, key=lambda t: (t[0], t[-1])), (name, k)
