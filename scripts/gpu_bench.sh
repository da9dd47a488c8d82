#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 25 gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2> gpurun_out/bench_nograph.err; echo "bench nograph rc=$?"
tail -c 2500 gpurun_out/bench_nograph.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
tail -c 800 gpurun_out/bench_ref.log; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
