#!/bin/bash
# round-2 GPU session 9 (1 GPU): the default bench line twice (VAE timing stability after the warm-up / trim-threshold change).
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 420 r02_bench_final3_N1.log python bench.py --gpus 1 --steps 20 --warmup 5; tail -n 1 gpurun_out/r02_bench_final3_N1.log | cut -c1-700
run 300 r02_bench_final4_N1.log python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-full-run; tail -n 1 gpurun_out/r02_bench_final4_N1.log | cut -c1-700
