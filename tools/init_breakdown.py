"""Where the pastar CLI's Phase 1 goes: the same calls through ctypes in a fresh process WITHOUT torch (so the CUDA
context is created by the library's first call, as in the CLI)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.perf_counter()
cudart = C.CDLL("libcudart.so.12")
cudart.cudaFree(None)
t1 = time.perf_counter()
import mpi_pastar_msa_b200.api as m   # numpy + ctypes only
from conftest import CASES
seqs = CASES["kinase"]
t2 = time.perf_counter()
w = m.gpu_weights(seqs)
t3 = time.perf_counter()
G = m.PastarGPU(seqs, weights=w.astype("int32"))
t4 = time.perf_counter()
G.build_pair_tables()
t5 = time.perf_counter()
r = G.search(table_capacity=0, batch_target=16384, want_rows=True)
t6 = time.perf_counter()
print("cuda context %.3f s | import %.3f | gpu_weights %.3f | ctx_create %.3f | pair tables %.3f | search %.3f (kernel %.1f ms) g=%d" % (
    t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, r["kernel_ms"], r["g"]))
