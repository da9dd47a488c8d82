#!/bin/bash
# development job on one GPU: GPU test suite, per-layer timings, bench.  Usage: gpurun -- bash scripts/gpu_job.sh <tag> [pytest-filter]
TAG=${1:-job}
FILTER=${2:-}
BENCH_FLAGS=${3:---no-cpu-baseline --no-sampling --no-torch-cuda-baseline --no-mechanics}
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 600 --timeout-method=thread"
if [ -n "$FILTER" ]; then
  timeout 1500 $PT tests -m gpu -x -k "$FILTER" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
else
  timeout 1500 $PT tests -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
fi
tail -n 15 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/layer_times.py > gpurun_out/${TAG}_layer_times.txt 2>&1; echo "layer times rc=$?"
head -n 30 gpurun_out/${TAG}_layer_times.txt
timeout 900 python bench.py $BENCH_FLAGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'), d.get('roofline', {}).get('frac'))
    print(json.dumps(d.get('kernel_time_breakdown_ms'), indent=0)[:1500])
except Exception as e:
    print('no json', e)
PY
tail -n 5 gpurun_out/${TAG}_bench.err
