"""python tools/cli_multi_gpu.py G : run the pastar CLI with -g G on kinase.fasta's sequences and check the optimal cost."""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CASES
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fa = os.path.join(tempfile.mkdtemp(), "kinase.fasta")
with open(fa, "w") as f:
    for i, s in enumerate(CASES["kinase"]):
        f.write(">Sequence %d\n%s\n" % (i + 1, s))
r = subprocess.run([os.path.join(ROOT, "mpi_pastar_msa_b200", "bin", "pastar"), "-g", str(G), "--batch", "16384", fa], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
out = r.stdout.decode()
print("\n".join(l for l in out.split("\n") if not re.match(r"^[A-Z\-]*$", l) or not l))
m = re.search(r"Final Score: .*g - (\d+)", out)
ok = r.returncode == 0 and m and int(m.group(1)) == 421546
print("CLI -g %d: %s" % (G, "OK" if ok else "FAILED (rc %d)" % r.returncode))
sys.exit(0 if ok else 1)
