"""bench.py's reference arm runs on host cores only (no GPU): its JSON line keeps the driver's contract.

The GPU arm's line is produced on the GPU box (profiles/r02_bench_n*.json); here only the CPU leg can run.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_ref():
    sys.path.insert(0, ROOT)
    from oracle import refio
    return refio.available()


def _run(env_extra=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    return r


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref is built by __graft_entry__.build() where /root/reference exists")
def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "node_expansions_per_sec" and d["unit"] == "expansions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    # SURVEY 8(d) micro-baselines, one core, reference code only
    mb = cb["micro"]
    assert mb["cores"] == 1 and mb["pair_align_gcups"] > 0 and mb["getneigh_parents_per_s"] > 0
    assert "S7" in d["config"]["workload"] or "N=7" in d["config"]["workload"]
    assert set(d["reference_build"]) >= {"mpicxx", "mpiexec", "boost_headers", "lz4_header", "buildable"}


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref is built by __graft_entry__.build() where /root/reference exists")
def test_reference_arm_other_ranks_do_no_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29571"})
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert not [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
