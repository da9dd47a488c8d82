#!/bin/bash
# bench.py (train step only) under explicit environment settings.  Usage: gpu_sweep2.sh <tag> "<ENV=V ENV2=V2>" ...
TAG=$1; shift
mkdir -p gpurun_out
FLAGS="--no-cpu-baseline --no-sampling --no-torch-cuda-baseline --no-mechanics --steps 30"
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py $FLAGS > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$i.json').read().strip().splitlines()[-1])
    k = d.get('kernel_time_breakdown_ms', {})
    print('$envs', 'ms/step', round(d['ms_per_step'], 4), 'conv', k.get('pidm_conv2d_tc_general', {}).get('ms'), 'wgrad', k.get('pidm_conv2d_wgrad_tc', {}).get('ms'))
except Exception as e:
    print('$envs', 'no json', e)
PY
done
