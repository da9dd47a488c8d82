#!/bin/bash
# bench.py (train step only) under several values of one tuning environment variable.
# Usage: gpu_sweep_env.sh <tag> <ENV_NAME> <value> [<value> ...]
TAG=$1; NAME=$2; shift 2
mkdir -p gpurun_out
FLAGS="--no-cpu-baseline --no-sampling --no-torch-cuda-baseline --no-mechanics --steps 30"
for v in "$@"; do
  env $NAME=$v timeout 600 python bench.py $FLAGS > gpurun_out/${TAG}_${NAME}_$v.json 2> gpurun_out/${TAG}_${NAME}_$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_${NAME}_$v.json').read().strip().splitlines()[-1])
    k = d.get('kernel_time_breakdown_ms', {})
    print('$NAME=$v', 'ms/step', round(d['ms_per_step'], 4), 'conv', k.get('pidm_conv2d_tc_general', {}).get('ms'), 'wgrad', k.get('pidm_conv2d_wgrad_tc', {}).get('ms'))
except Exception as e:
    print('$NAME=$v', 'no json', e)
PY
done
