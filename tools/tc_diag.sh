#!/bin/bash
# run every diagnostic case in its own process so that a trap does not poison the rest
for c in one_box_c32 one_box_c64 one_box_c256 halo_c64 ragged_c96 cfg1; do
  echo "=== $c"
  timeout 120 python tools/tc_diag.py $c 2>&1 | tail -30
done
