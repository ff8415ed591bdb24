"""Wall time of the pastar CLI on kinase.fasta (process start to exit), three runs: python tools/cli_kinase_time.py [-g G]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASES
fa = "/tmp/kinase.fasta"
open(fa, "w").write("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(CASES["kinase"])))
for _ in range(3):
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "mpi_pastar_msa_b200", "bin", "pastar")] + sys.argv[1:] + [fa], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    dt = time.perf_counter() - t0
    lines = [l for l in r.stdout.decode().splitlines() if "Phase" in l or "Final" in l]
    print("wall %.3f s rc %d | %s" % (dt, r.returncode, " | ".join(lines)), flush=True)
