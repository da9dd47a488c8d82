#!/bin/bash
# quick iteration: targeted tests ($1 = pytest -k expression), then the e2e suite and the bench
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 240 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_ops.py -k "$1" > gpurun_out/quick_ops.log 2>&1; echo "quick ops rc=$?"
grep -E "^E  |FAILED|passed|failed" gpurun_out/quick_ops.log | head -40
timeout 900 $PT tests/test_gpu_e2e.py > gpurun_out/e2e_tc.log 2>&1; echo "e2e rc=$?"
grep -E "^E  |FAILED|passed|failed" gpurun_out/e2e_tc.log | head -20
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    print('roofline', d['roofline'])
    print('residual', {k: d['roofline_residual'][k] for k in ('achieved', 'frac')}, d['roofline_residual']['fused_loss_grad_variant'])
    for k, v in d['kernel_time_breakdown_ms'].items():
        print(f"  {k:32s} {v['ms']:8.3f} ms  calls {v['calls']:4d}  tflops {v['tflops']}")
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('no bench json', e)
PY
