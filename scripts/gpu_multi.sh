#!/bin/bash
# multi-GPU run: $1 = number of GPUs (short timeouts: a hang costs N x the box time)
mkdir -p gpurun_out
N=$1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-sampling > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -n 4 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e:
    print('no json', e)
PY
