#!/bin/sh
# One `ncu --set full` capture each of the strict and the fast fused-step kernel at 8192^2 (run on the GPU box under
# gpurun), plus the launch list of the bench command.  Writes reports and summaries into gpurun_out/<tag>_*; back in
# the build container `tools/ncu_summarise.py <tag>` turns them into profiles/<tag>_* and re-stamps
# profiles/roofline_traffic.json with the sha of the CUDA sources they were captured from.
#   usage (on the box): tools/ncu_capture.sh r2a
tag=${1:-r2}
out=gpurun_out
for mode in strict fast; do
  ncu --set full --import-source on --clock-control none -k regex:k_fused_step -s 6 -c 1 -f -o $out/${tag}_$mode \
      python tools/quick_perf.py four_quadrant 8192 8192 2 $mode > $out/${tag}_ncu_$mode.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_8192.csv \
    python bench.py --steps 3 --warmup 3 --no-fast --no-cpu-baseline --no-configs --no-parity > $out/${tag}_ncu_bench.log 2>&1
