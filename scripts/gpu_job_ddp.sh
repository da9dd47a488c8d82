#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_job_ddp.sh N tag
N=${1:-2}; TAG=${2:-ddp}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  timeout 240 $TR scripts/check_ddp.py > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; grep "^\[\|DDP_CHECK" gpurun_out/${TAG}_check.log; tail -3 gpurun_out/${TAG}_check.log
fi
FL="--no-cpu-baseline --no-sampling --no-torch-cuda-baseline --no-mechanics"
for MODE in 0 1; do
  PIDM_BUCKET_AR=$MODE timeout 240 $TR bench.py --gpus $N --steps 30 --warmup 5 $FL > gpurun_out/${TAG}_bench_ar$MODE.json 2> gpurun_out/${TAG}_bench_ar$MODE.err; echo "bench AR=$MODE rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_ar$MODE.json').read().strip().splitlines()[-1])
    print('AR=$MODE', {k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step')}, d.get('e2e', {}).get('value'))
except Exception as e:
    print('no json', e)
PY
  tail -2 gpurun_out/${TAG}_bench_ar$MODE.err
done
