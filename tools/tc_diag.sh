#!/bin/bash
# run every diagnostic case in its own process so that a trap does not poison the rest
SPLIT=${1:-tf32}
for c in one_box_c64 one_box_c256 halo_c64 cfg1; do
  echo "=== $c $SPLIT"
  timeout 120 python tools/tc_diag.py $c $SPLIT 2>&1 | tail -30
done
