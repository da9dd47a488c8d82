#!/bin/bash
# A/B of one tuning environment variable: GPU tests, then per-layer timings and a short bench under each value.
# Usage: gpu_job_ab.sh <tag> <pytest filter> <ENV_NAME> <value> [<value> ...]
TAG=$1; FILTER=$2; NAME=$3; shift 3
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 600 --timeout-method=thread"
timeout 1500 $PT tests -m gpu -x -k "$FILTER" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/${TAG}_pytest.log
for v in "$@"; do
  env $NAME=$v timeout 300 python scripts/layer_times.py > gpurun_out/${TAG}_layer_times_$v.txt 2>&1; echo "layer times $NAME=$v rc=$?"
  head -n 8 gpurun_out/${TAG}_layer_times_$v.txt
done
bash scripts/gpu_sweep_env.sh $TAG $NAME "$@"
