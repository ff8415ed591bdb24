"""Summarise an .ncu-rep (read here, no GPU): key counters + the source lines with the most stall samples.
usage: python tools/ncu_summary.py report.ncu-rep [n_lines] [byinst]   (byinst: rank the lines by instructions executed)"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 14
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("launch__grid_size"))
    for i, h in enumerate(hdr):
        if h in WANT:
            print("  %-70s %-14s %s" % (h, units[i], r[i]))
    st = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)
    print("  stalls per issue:", ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for v, h in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
cur = ""
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        ln = int(d["Line No"]); smp = int(d.get("# Samples") or 0); ins = int(d.get("Instructions Executed") or 0)
    except ValueError:
        continue
    a = agg[(cur, ln)]
    a[0] += smp; a[1] += ins; a[2] = d.get("Source", "").strip()[:110]
    tot += smp
print("total samples", tot)
byinst = len(sys.argv) > 3 and sys.argv[3] == "byinst"
if byinst:
    print("total instructions", sum(a[1] for a in agg.values()))
for (f, ln), (smp, ins, text) in sorted(agg.items(), key=lambda x: -x[1][1 if byinst else 0])[:nl]:
    print("  %6d %5.1f%% %-22s:%-5d inst=%-9d %s" % (smp, 100.0 * smp / max(1, tot), f, ln, ins, text))
