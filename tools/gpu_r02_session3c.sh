#!/bin/bash
# round-2 GPU session 3c (1 GPU, short): conv parity after the producer-loop rewrite, per-shape table, cfg4 at N=1 (few steps).
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 240 r02_pytest_convs.log python -m pytest tests/test_gpu_ops.py tests/test_gpu_bench_shapes.py tests/test_gpu_unet.py -q -m gpu -k "conv or unet or lambda" -s --timeout 150; tail -n 8 gpurun_out/r02_pytest_convs.log
SUPIR_BENCH_DUMP_SHAPES=gpurun_out/r02_gemm_shapes_B98.json run 300 r02_bench_cfg3_N1_c.log python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-run; tail -n 2 gpurun_out/r02_bench_cfg3_N1_c.log | cut -c1-400
run 420 r02_bench_cfg4_N1.log python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-full-run; tail -n 2 gpurun_out/r02_bench_cfg4_N1.log | cut -c1-900
