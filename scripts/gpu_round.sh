#!/bin/bash
# One GPU session: parity tests (CUDA-core conv path first, then the tcgen05 path), smoke, a short bench.
# Every stage has its own timeout so that a hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q -p no:cacheprovider --timeout 240 --timeout-method=thread"
PIDM_DISABLE_TC=1 timeout 900 $PT tests/test_gpu_ops.py -k "not tcgen05" > gpurun_out/ops_simt.log 2>&1; echo "ops_simt rc=$?"
PIDM_DISABLE_TC=1 timeout 900 $PT tests/test_gpu_e2e.py -k "not tcgen05 and not smoke" > gpurun_out/e2e_simt.log 2>&1; echo "e2e_simt rc=$?"
timeout 600 $PT tests/test_gpu_ops.py -k "tcgen05" > gpurun_out/ops_tc.log 2>&1; echo "ops_tc rc=$?"
timeout 900 $PT tests/test_gpu_e2e.py > gpurun_out/e2e_tc.log 2>&1; echo "e2e_tc rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 3 gpurun_out/ops_simt.log gpurun_out/e2e_simt.log gpurun_out/ops_tc.log gpurun_out/e2e_tc.log
tail -c 1500 gpurun_out/bench.log
