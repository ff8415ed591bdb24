#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): one --set full capture of every hot kernel in steady state + the launch
# list of the search round.  Reports land in gpurun_out/; tools/ncu_summary.py turns them into profiles/r02_*.txt.
set -x
bash tools/ncu_round.sh r02 "expand_probe_kernel insert_kernel claim_kernel"
ncu --set full --clock-control none --import-source on -k regex:pair_dp_linear -s 2 -c 1 -f -o gpurun_out/prof_r02_pair_dp_s7 python tools/dp_once.py > gpurun_out/ncu_r02_dp_s7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_dp_linear -s 2 -c 1 -f -o gpurun_out/prof_r02_pair_dp_s8 python tools/dp_once.py s8 > gpurun_out/ncu_r02_dp_s8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_expand_batch_s7 python tools/expand_once.py s7 > gpurun_out/ncu_r02_xb_s7.log 2>&1
ls -la gpurun_out/prof_r02_*
# summaries are made here (the reports are ~45 MB each: too big to bring back); only text leaves the box
for k in expand_probe_kernel insert_kernel claim_kernel pair_dp_s7 pair_dp_s8 expand_batch_s7; do
  python tools/ncu_summary.py gpurun_out/prof_r02_$k.ncu-rep 24 > gpurun_out/r02_ncu_$k.txt 2>&1
done
rm -f gpurun_out/prof_r02_*.ncu-rep
