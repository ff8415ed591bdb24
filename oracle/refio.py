"""TEST INFRASTRUCTURE: run oracle/_ref/pastar_ref (the unmodified reference
arithmetic + ref_driver.cpp) as a subprocess and parse its binary dumps."""
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref", "pastar_ref")


def available():
    return os.path.exists(REF_BIN)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">s%d\n%s\n" % (i, s))


def _run(args, timeout=3600):
    r = subprocess.run([REF_BIN] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("pastar_ref %s failed rc=%d: %s" % (args, r.returncode, r.stderr.decode()[-400:]))
    return r.stdout.decode()


def dump(seqs):
    """-> dict(n, lens, cost[90,90], weights[n,n] float32, tables[list of int32 2-D])"""
    with tempfile.TemporaryDirectory() as d:
        fa, out = os.path.join(d, "in.fasta"), os.path.join(d, "out.bin")
        write_fasta(fa, seqs)
        _run(["dump", fa, out])
        b = open(out, "rb").read()
    assert b[:4] == b"PGD1"
    o = 4
    (n,) = struct.unpack_from("<i", b, o)
    o += 4
    lens = list(struct.unpack_from("<%di" % n, b, o))
    o += 4 * n
    cost = np.frombuffer(b, dtype="<i4", count=8100, offset=o).reshape(90, 90).copy()
    o += 4 * 8100
    w = np.frombuffer(b, dtype="<f4", count=n * n, offset=o).reshape(n, n).copy()
    o += 4 * n * n
    tables = []
    for _ in range(n * (n - 1) // 2):
        r, c = struct.unpack_from("<ii", b, o)
        o += 8
        tables.append(np.frombuffer(b, dtype="<i4", count=r * c, offset=o).reshape(r, c).copy())
        o += 4 * r * c
    return {"n": n, "lens": lens, "cost": cost, "weights": w, "tables": tables}


def neigh(seqs, parents_pos, parents_g, parents_parenti, vec_size=1, hash_type="FZORDER", shift=12):
    """Node<N>::getNeigh for every parent. -> list of (f_parent, structured successor array)."""
    from .oracle import succ_dtype
    n = len(seqs)
    pos = np.ascontiguousarray(parents_pos, dtype="<u2").reshape(-1, n)
    k = pos.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fa, pin, out = os.path.join(d, "in.fasta"), os.path.join(d, "p.bin"), os.path.join(d, "out.bin")
        write_fasta(fa, seqs)
        with open(pin, "wb") as f:
            f.write(struct.pack("<i", k))
            for i in range(k):
                f.write(pos[i].tobytes())
                f.write(struct.pack("<ii", int(parents_g[i]), int(parents_parenti[i])))
        _run(["neigh", fa, pin, out, vec_size, hash_type, shift])
        b = open(out, "rb").read()
    assert b[:4] == b"PGN1"
    n2, k2 = struct.unpack_from("<ii", b, 4)
    assert n2 == n and k2 == k
    o = 12
    rec = np.dtype([("pos", "<u2", (n,)), ("f", "<i4"), ("g", "<i4"), ("parenti", "<i4"), ("owner", "<i4")])
    res = []
    for _ in range(k):
        fpar, cnt = struct.unpack_from("<ii", b, o)
        o += 8
        a = np.frombuffer(b, dtype=rec, count=cnt, offset=o)
        o += rec.itemsize * cnt
        s = np.zeros(cnt, dtype=succ_dtype(n))
        for name in ("pos", "f", "g", "parenti"):
            s[name] = a[name]
        s["owner"] = a["owner"].astype(np.uint32)
        res.append((fpar, s))
    return res


def owner(n, coords, vec_size, hash_type="FZORDER", shift=12):
    """Coord<N>::get_id for each coord (needs any n-sequence problem loaded: uses dummies)."""
    pos = np.ascontiguousarray(coords, dtype="<u2").reshape(-1, n)
    with tempfile.TemporaryDirectory() as d:
        fa, pin, out = os.path.join(d, "in.fasta"), os.path.join(d, "p.bin"), os.path.join(d, "out.bin")
        write_fasta(fa, ["AC"] * n)
        with open(pin, "wb") as f:
            f.write(struct.pack("<i", pos.shape[0]))
            f.write(pos.tobytes())
        _run(["owner", fa, pin, out, vec_size, hash_type, shift])
        b = open(out, "rb").read()
    assert b[:4] == b"PGO1"
    (k,) = struct.unpack_from("<i", b, 4)
    return np.frombuffer(b, dtype="<i4", count=k, offset=8).astype(np.uint32)


def _json_cmd(args, timeout=3600):
    out = _run(args, timeout)
    for line in out.splitlines():
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("no JSON from pastar_ref: " + out[-300:])


def astar(seqs_or_path, budget=0, timeout=3600):
    return _search(["astar"], seqs_or_path, [budget], timeout)


def pastar(seqs_or_path, threads, budget=0, hash_type="FZORDER", shift=12, timeout=3600, warm_pops=0):
    return _search(["pastar"], seqs_or_path, [threads, budget, hash_type, shift, warm_pops], timeout)


def micro(seqs_or_path, seconds=2.0, timeout=600):
    return _search(["micro"], seqs_or_path, [seconds], timeout)


def _search(cmd, seqs_or_path, tail, timeout):
    if isinstance(seqs_or_path, str):
        return _json_cmd(cmd + [seqs_or_path] + tail, timeout)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fasta")
        write_fasta(fa, seqs_or_path)
        return _json_cmd(cmd + [fa] + tail, timeout)
