#!/bin/sh
# A/B timing of library variants, `arithmetic=fast` (development aid). usage: tools/ab_run_fast.sh out.txt variant...
out=$1; shift
: > "$out"
for v in "$@"; do
  echo "== $v" >> "$out"
  E2D_LIB_PATH=build/variants/$v/libeuler2d_b200.so python tools/quick_perf.py four_quadrant 8192 8192 20 fast 2>&1 | tail -2 >> "$out"
done
