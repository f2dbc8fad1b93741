#!/bin/sh
run() { printf "%-22s taper=%s guide=%-4s min_rows=%-3s " "$1 $2x$3" "$4" "$5" "$6"; E2D_SEG_TAPER=$4 E2D_SEG_GUIDE=$5 E2D_SEG_MIN_ROWS=$6 python tools/quick_perf.py $1 $2 $3 $7 ${8:-strict} 2>&1 | tail -1 | sed 's/.*-> //; s/ (.*//'; }
for g in 1.0 1.25 1.5 1.75; do for m in 32 40 48; do run four_quadrant 8192 8192 1 $g $m 20; done; done
for g in 1.0 1.25 1.5; do for m in 32 48; do run four_quadrant 16384 2048 1 $g $m 40; run four_quadrant 32768 4096 1 $g $m 10; done; done
