#!/bin/bash
# Round-end evidence run on one GPU (tag r02): GPU test suite, smoke, bench (both arms), ncu launch list of one eager step,
# full ncu captures of the kernels that dominate the step -- the convolution instantiations are picked BY TIME from the
# launch list of round 1 (<32,64>, <32,32>, <64,64>), not by launch index -- and per-layer graph timings.
# Numbers printed by runs under ncu are never bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 600 --timeout-method=thread"
timeout 1200 $PT tests -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sampling --no-mechanics --no-torch-cuda-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches_run.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_step_traffic.json > gpurun_out/${TAG}_launch_summary.txt 2>&1
head -n 12 gpurun_out/${TAG}_launch_summary.txt
cap() {   # cap <file tag> <demangled-name regex> <skip> <count>
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f \
      -o gpurun_out/${TAG}_prof_$1 $B > gpurun_out/${TAG}_prof_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap conv_32_64 'conv_tc_kernel<.int.32, .int.64>' 250 2
cap conv_32_32 'conv_tc_kernel<.int.32, .int.32>' 90 2
cap conv_64_64 'conv_tc_kernel<.int.64, .int.64>' 55 1
cap wgrad3 'wgrad3_kernel' 60 2
cap gn_bwd_piece 'gn_bwd_piece_kernel' 120 3
cap gn_apply 'gn_apply_kernel' 120 1
cap laf_bwd 'laf_bwd_kernel' 9 1
cap laf_out 'laf_out_kernel' 9 1
cap darcy_grad 'darcy_grad_kernel' 3 1
cap adam 'adam_ema_kernel' 3 1
timeout 300 python scripts/layer_times.py > gpurun_out/${TAG}_layer_times.txt 2>&1; echo "layer times rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench.json', 'gpurun_out/${TAG}_bench_reference.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ('impl', 'value', 'ms_per_step', 'gpu_launches')}, d.get('e2e', {}).get('value'), d.get('clocks'))
    except Exception as e:
        print(f, 'no json', e)
PY
