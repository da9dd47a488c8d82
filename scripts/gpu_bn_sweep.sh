#!/bin/bash
# n-tile width sweep of the tensor-core conv on the small-M layers (debug aid, PIDM_TC_BN override)
for shape in "32 8 256 256 3" "32 8 128 128 3" "32 16 128 128 3" "32 16 64 64 3" "32 32 64 64 3" "32 8 512 128 3" "32 8 256 256 1"; do
  for bn in 32 64 128 256; do
    echo -n "BN=$bn  "; PIDM_TC_BN=$bn timeout 120 python scripts/trace_conv.py $shape q 2>&1 | tail -1
  done
done
