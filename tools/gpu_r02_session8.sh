#!/bin/bash
# round-2 GPU session 8 (1 GPU): the batched embedding projection on tensor cores — network parity incl. a batch above 16 rows,
# then the default bench line.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 400 r02_pytest_unet_b.log python -m pytest tests/test_gpu_unet.py tests/test_gpu_vae_sampler.py tests/test_gpu_model.py -q -m gpu -s --timeout 200; tail -n 12 gpurun_out/r02_pytest_unet_b.log | cut -c1-300
run 420 r02_bench_final2_N1.log python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline; tail -n 1 gpurun_out/r02_bench_final2_N1.log | cut -c1-1200
