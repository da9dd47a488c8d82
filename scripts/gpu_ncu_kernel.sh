#!/bin/bash
# full ncu capture of one kernel family: $1 = kernel regex, $2 = launches to skip, $3 = launches to capture
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sampling"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/prof_$1 $B > gpurun_out/prof_$1.log 2>&1; echo "ncu $1 rc=$?"
