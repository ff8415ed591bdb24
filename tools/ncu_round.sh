#!/bin/bash
# ncu captures of one steady-state launch of each kernel of the search round (S7, batch 2^20, 16 GiB table), plus the
# launch list of 40 steady-state launches.  --cache-control none: the rounds run back to back with a warm L2 (the 64 MB block
# directory and the pairwise tables stay resident between kernels); ncu's default flush before every kernel makes the expand
# kernel re-fetch them (465 us instead of 303 us).   usage (under gpurun): bash tools/ncu_round.sh <tag> [kernel-regex ...]
tag=$1; shift
kernels=${@:-"insert_kernel claim_kernel expand_probe_kernel"}
cmd="python tools/search_once.py s7 1048576 60000000 1073741824"
$cmd > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'select_kernel|claim_kernel|expand_probe_kernel|insert_kernel' -s 1600 -c 40 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_launches_$tag.log 2>&1
for k in $kernels; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 420 -c 1 -f -o gpurun_out/prof_${tag}_$k $cmd > gpurun_out/ncu_${tag}_$k.log 2>&1
done
tail -2 gpurun_out/plain_$tag.log
