#!/bin/bash
# ncu captures of the routing kernels (one GPU): for every workload the plain run first, then
# `--set full` on two launches after warm-up; reports land in gpurun_out/r02/ (read them on the
# CPU box with scripts/ncu_summary.py).  usage: gpurun -- bash scripts/ncu_capture.sh [workloads...]
out=gpurun_out/r02; mkdir -p $out
for w in ${@:-kw_ga kw dw_full}; do
  python scripts/ncu_target.py 4096 $w 12 > $out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name regex:channels_ \
      --launch-skip 6 --launch-count 3 -f \
      -o $out/prof_$w python scripts/ncu_target.py 4096 $w 12 > $out/ncu_$w.log 2>&1
  tail -2 $out/ncu_$w.log
done
# launch list of the bench command itself
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity"
$B > $out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv $B > $out/ncu_bench.log 2>&1
tail -2 $out/ncu_bench.log; ls -la $out | tail -20
