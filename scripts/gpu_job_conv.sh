#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 120 --timeout-method=thread"
PIDM_TC_CPASYNC=1 timeout 300 $PT tests/test_gpu_ops.py -m gpu -k "conv or wgrad" -x > gpurun_out/jc_pytest1.log 2>&1; echo "conv tests (cpasync) rc=$?"; tail -3 gpurun_out/jc_pytest1.log | cut -c1-200
for CP in 1 0; do
  echo "== cpasync=$CP"; PIDM_TC_CPASYNC=$CP timeout 120 python scripts/trace_conv.py 32 8 256 256 3 2>&1 | sed -n 1,3p
  PIDM_TC_CPASYNC=$CP timeout 120 python scripts/trace_conv.py 32 64 32 32 3 x 2>&1 | head -1
  PIDM_TC_CPASYNC=$CP timeout 120 python scripts/trace_conv.py 32 16 128 128 3 x 2>&1 | head -1
  PIDM_TC_CPASYNC=$CP timeout 120 python scripts/trace_conv.py 32 32 64 64 3 x 2>&1 | head -1
done
