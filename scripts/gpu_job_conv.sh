#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 60 --timeout-method=thread"
timeout 200 $PT tests/test_gpu_ops.py -m gpu -k "conv or wgrad" > gpurun_out/jc_pytest1.log 2>&1; echo "conv tests (rg2) rc=$?"; grep -E "passed|failed|FAILED|rel" gpurun_out/jc_pytest1.log | head -12 | cut -c1-220
for RG in 2 1; do
  echo "== PIDM_TC_RG=$RG"
  PIDM_TC_RG=$RG timeout 60 python scripts/trace_conv.py 32 64 32 32 3 x 2>&1 | head -1
  PIDM_TC_RG=$RG timeout 60 python scripts/trace_conv.py 32 32 64 64 3 x 2>&1 | head -1
  PIDM_TC_RG=$RG timeout 60 python scripts/trace_conv.py 32 16 128 128 3 x 2>&1 | head -1
done
