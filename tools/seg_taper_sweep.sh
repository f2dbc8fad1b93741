#!/bin/sh
# Development aid: tapered segments (taper_segments, e2d_kernels.cu) against uniform ones, and the two tuning knobs.
run() { printf "%-22s taper=%s guide=%-4s min_rows=%-3s " "$1 $2x$3" "$4" "$5" "$6"; E2D_SEG_TAPER=$4 E2D_SEG_GUIDE=$5 E2D_SEG_MIN_ROWS=$6 python tools/quick_perf.py $1 $2 $3 $7 ${8:-strict} 2>&1 | tail -1 | sed 's/.*-> //; s/ (.*//'; }
run four_quadrant 8192 8192 0 2 40 20
for g in 1.5 2 3; do for m in 24 40 64; do run four_quadrant 8192 8192 1 $g $m 20; done; done
run four_quadrant 8192 8192 0 2 40 20
for shape in "16384 2048 40" "32768 4096 10" "4096 4096 40" "16384 16384 6" "2048 8192 40"; do
  set -- $shape
  run four_quadrant $1 $2 0 2 40 $3
  run four_quadrant $1 $2 1 2 40 $3
  run four_quadrant $1 $2 1 1.5 40 $3
  run four_quadrant $1 $2 1 3 24 $3
done
run blast 1024 1536 0 2 40 400
run blast 1024 1536 1 2 40 400
run four_quadrant 8192 8192 0 2 40 20 fast
run four_quadrant 8192 8192 1 2 40 20 fast
run four_quadrant 8192 8192 1 1.5 40 20 fast
