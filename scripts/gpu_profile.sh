#!/bin/bash
# ncu evidence for the step: (1) launch list with per-launch device time, (2) full captures of the top kernels.
# Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_run.log 2>&1; echo "launch list rc=$?"
for K in conv_tc_kernel conv_wgrad_simt_kernel darcy_kernel la_bwd_pixel_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 3 -f -o gpurun_out/prof_$K $B > gpurun_out/prof_$K.log 2>&1; echo "ncu $K rc=$?"
done
ls -la gpurun_out | head -30
