#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 120 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_ops.py -m gpu -k "conv or wgrad" > gpurun_out/jc_pytest1.log 2>&1; echo "conv tests rc=$?"; tail -4 gpurun_out/jc_pytest1.log
timeout 600 $PT tests/test_gpu_parity_bench_path.py tests/test_gpu_e2e.py -m gpu > gpurun_out/jc_pytest2.log 2>&1; echo "parity tests rc=$?"; tail -6 gpurun_out/jc_pytest2.log
for SP in 1 0; do
  PIDM_TC_SPLIT=$SP timeout 200 python scripts/layer_times.py conv2d_tc_general > gpurun_out/jc_lt_split$SP.txt 2>&1
  echo "== split=$SP"; head -2 gpurun_out/jc_lt_split$SP.txt; grep " 8, 8, \| 16, 16, 128, 8" gpurun_out/jc_lt_split$SP.txt | head -14
done
