#!/bin/bash
# full ncu captures of single kernels of the eager training step.  Each argument: "<file tag>|<demangled-name regex>|<skip>|<count>[|ENV=VALUE]"
TAG=${TAG:-ncu}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sampling --no-mechanics --no-torch-cuda-baseline"
for spec in "$@"; do
  IFS='|' read -r name regex skip count envs <<< "$spec"
  env $envs timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s $skip -c $count -f \
      -o gpurun_out/${TAG}_prof_$name $B > gpurun_out/${TAG}_prof_$name.log 2>&1; echo "ncu $name rc=$?"; ls -la gpurun_out/${TAG}_prof_$name.ncu-rep
done
