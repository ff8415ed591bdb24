"""Generate tests/golden/reference_golden.npz from the UNMODIFIED reference (oracle/_ref/pastar_ref).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures pin: Altschul weights (float32 bit patterns), pairwise-table int64 sums + corner cells,
h(start), Node::getNeigh records (pos, f, g, parenti, owner) for seeded parents under two hash
configurations, Coord::get_id samples for all four hashes, and the optimal cost g* where the reference's
serial A* (AStar.cpp:53-104 semantics) terminates in about a minute.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASES, S7, random_parents  # noqa: E402
from oracle import refio as R  # noqa: E402

HASH_CFGS = [(1, "FZORDER", 12), (5, "FZORDER", 3)]
ASTAR_OK = {"test", "test2", "PF08184", "rnd4x60", "fam6x80", "fam3x300", "fam5x60", "fam4x150", "fam7x30", "fam8x20",
            "fam9x30", "fam10x20"}


def main():
    assert R.available(), "build oracle/_ref first"
    out = {}
    cases = dict(CASES)
    cases["S7"] = S7()
    for name, seqs in cases.items():
        n = len(seqs)
        d = R.dump(seqs)
        out[name + "/seqs"] = np.array(seqs)
        out[name + "/weights_f32"] = d["weights"]
        out[name + "/table_sums"] = np.array([int(t.astype(np.int64).sum()) for t in d["tables"]], dtype=np.int64)
        out[name + "/table_corner"] = np.array([[t[0, 0], t[0, -1], t[-1, 0], t[t.shape[0] // 2, t.shape[1] // 3]] for t in d["tables"]],
                                               dtype=np.int32)
        if name == "test":
            out["cost_table"] = d["cost"]
        if n <= 10:
            k = 8 if n <= 8 else 2
            pos, g, par = random_parents(seqs, k, 7)
            out[name + "/parents_pos"], out[name + "/parents_g"], out[name + "/parents_par"] = pos, g, par
            for ci, (vs, ht, sh) in enumerate(HASH_CFGS):
                res = R.neigh(seqs, pos, g, par, vs, ht, sh)
                out[name + "/neigh%d_fpar" % ci] = np.array([r[0] for r in res], dtype=np.int32)
                out[name + "/neigh%d_count" % ci] = np.array([len(r[1]) for r in res], dtype=np.int32)
                cat = np.concatenate([r[1] for r in res])
                for fld in ("pos", "f", "g", "parenti", "owner"):
                    out[name + "/neigh%d_%s" % (ci, fld)] = cat[fld]
        if name in ASTAR_OK:
            a = R.astar(seqs)
            assert a["finished"] == 1
            out[name + "/astar"] = np.array([a["g"], a["expansions"], a["generated"]], dtype=np.int64)
            print(name, "g*", a["g"], "exp", a["expansions"], flush=True)
    # kinase: 6-7 minutes of serial A*; value re-derived with `oracle/_ref/pastar_ref astar kinase.fasta`
    out["kinase/astar"] = np.array([421546, 4497279, 139409688], dtype=np.int64)
    # owner hashes
    rng = np.random.default_rng(9)
    for n in (3, 5, 7, 8, 10, 16):
        co = rng.integers(0, 65536, (64, n)).astype(np.uint16)
        co[:32] %= 1024
        out["owner/%d/coords" % n] = co
        for ht in ("FZORDER", "PZORDER", "FSUM", "PSUM"):
            for sh in (0, 5, 12, 21):
                for size in (1, 3, 8, 64):
                    out["owner/%d/%s/%d/%d" % (n, ht, sh, size)] = R.owner(n, co, size, ht, sh)
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), os.path.getsize(os.path.join(HERE, "reference_golden.npz")))


if __name__ == "__main__":
    main()
