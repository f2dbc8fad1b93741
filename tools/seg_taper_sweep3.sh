#!/bin/sh
# Development aid: from which grid size on the tapered segments pay (E2D_SEG_TAPER_FAIR: fair share of rows per slot, in minimal segments)
run() { printf "%-24s taper=%s min_rows=%-3s fair=%-4s " "$1 $2x$3" "$4" "$5" "$6"; E2D_SEG_TAPER=$4 E2D_SEG_MIN_ROWS=$5 E2D_SEG_TAPER_FAIR=$6 python tools/quick_perf.py $1 $2 $3 $7 strict 2>&1 | tail -1 | sed 's/.*-> //; s/ (.*//'; }
for shape in "4096 4096 60" "2048 8192 60" "6144 6144 30" "3072 3072 100" "2048 2048 200" "1024 1536 400" "8192 1024 100" "1536 1536 300"; do
  set -- $shape
  run four_quadrant $1 $2 0 48 4 $3
  run four_quadrant $1 $2 1 48 0.5 $3
  run four_quadrant $1 $2 1 32 0.5 $3
  run four_quadrant $1 $2 1 24 0.5 $3
done
