#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_parity_bench_path.py -m gpu > gpurun_out/j3_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/j3_pytest.log | tail -20
for CL in 0 1 2 4 8; do
  PIDM_GN_CL=$CL timeout 200 python scripts/layer_times.py groupnorm_silu_bwd > gpurun_out/j3_gn_cl$CL.txt 2>&1
  echo "== CL=$CL"; grep "groupnorm_silu_bwd  " gpurun_out/j3_gn_cl$CL.txt | head -12
done
