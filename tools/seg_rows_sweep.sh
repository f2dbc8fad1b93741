#!/bin/sh
# Development aid: throughput of the strict step at 8192^2 against the forced segment length (E2D_SEG_ROWS).
for r in 0 40 78 103 124 155 205 316 631 1366; do
  echo "== E2D_SEG_ROWS=$r"
  E2D_SEG_ROWS=$r python tools/quick_perf.py four_quadrant 8192 8192 20 strict 2>&1 | tail -2
done
