#!/bin/bash
# full ncu captures of the three tensor-core convolution instantiations that dominate the step BY TIME (launch list:
# <32,64> 1.05 ms, <32,32> 0.40 ms, <64,64> 0.19 ms per step).  ncu prints template arguments as "(int)32".
TAG=${1:-r02}
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sampling --no-mechanics --no-torch-cuda-baseline"
cap() {   # cap <file tag> <demangled-name regex> <skip> <count>
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f \
      -o gpurun_out/${TAG}_prof_$1 $B > gpurun_out/${TAG}_prof_$1.log 2>&1; echo "ncu $1 rc=$?"; ls -la gpurun_out/${TAG}_prof_$1.ncu-rep
}
cap conv_32_64 'conv_tc_kernel<.int.32, .int.64>' 250 2
cap conv_32_32 'conv_tc_kernel<.int.32, .int.32>' 90 2
cap conv_64_64 'conv_tc_kernel<.int.64, .int.64>' 55 1
