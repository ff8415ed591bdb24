"""python tools/print_bench.py bench.json : the few numbers of a bench.py line one looks at first."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("%d GPU(s): %.1f M expansions/s, %.3f ms/round, e2e %.1f M/s" % (d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], ((d.get("e2e") or {}).get("value") or 0) / 1e6))
r = d.get("roofline")
if r:
    print("roofline: %s frac %.3f of %s GB/s, whole round %.3f, traffic %s" % (r["kernel"], r["frac"], r["peak"], r.get("whole_round_frac", 0), r.get("traffic")))
    if "kernels" in r:
        print("kernels (us):", {k: round(v["ms_per_step"] * 1e3, 1) for k, v in r["kernels"].items()})
x = d.get("extra", {})
if "pair_dp" in x:
    print("pair DP: S7 %.1f GCUPS" % x["pair_dp"]["gcups"], "S8 %.1f GCUPS" % x.get("s8", {}).get("pair_dp", {}).get("gcups", 0),
          "| expand only: S7 %.2f of HBM, S8 %.2f" % (x["expand_only"]["frac_of_hbm"], x.get("s8", {}).get("expand_only", {}).get("frac_of_hbm", 0)))
if "rank0_kernel_ms_per_step" in x:
    print("rank 0 kernels (ms):", x["rank0_kernel_ms_per_step"], "NVLink bytes/round/GPU", x.get("nvlink_bytes_per_step_per_gpu"))
c = d.get("cpu_baseline")
if c:
    print("cpu baseline: %.0f expansions/s on %d threads (%s)" % (c["value"] or 0, c["cores"], c["kind"]))
