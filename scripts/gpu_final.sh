#!/bin/bash
# round-end evidence run on one GPU: full GPU test suite, smoke, bench (both arms), ncu launch list, full ncu captures of
# the top kernels, per-layer graph timings.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 --timeout-method=thread"
timeout 900 $PT tests -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "bench reference rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sampling"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_launches_run.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/final_launches.csv gpurun_out/final_step_traffic.json > gpurun_out/final_launch_summary.txt 2>&1
for K in conv_tc_kernel wgrad3_kernel gn_bwd_cluster_kernel laf_bwd_kernel darcy_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f -o gpurun_out/final_prof_$K $B > gpurun_out/final_prof_$K.log 2>&1; echo "ncu $K rc=$?"
done
timeout 300 python scripts/layer_times.py > gpurun_out/final_layer_times.txt 2>&1; echo "layer times rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/final_bench.json', 'gpurun_out/final_bench_reference.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ('impl', 'value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'), d.get('clocks'))
    except Exception as e:
        print(f, 'no json', e)
PY
