#!/bin/bash
# round-2 GPU session 3 (1 GPU): evidence. ncu launch list of one timed step + full captures, the reference's own GPU path
# (torch eager bf16) beside ours, the other BASELINE configurations at N=1, and the default bench line with its CPU leg.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 60 r02_selftest_attnquick3.log tools/selftest attnquick || { tail -n 20 gpurun_out/r02_selftest_attnquick3.log; echo "HANG GUARD FAILED"; exit 1; }
run 200 r02_pytest_1head.log python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_vae_sampler.py -q -m gpu -k "single_head or vae" -s --timeout 120; tail -n 12 gpurun_out/r02_pytest_1head.log
run 120 r02_selftest_attnstagger.log tools/selftest attnstagger; cat gpurun_out/r02_selftest_attnstagger.log
SUPIR_BENCH_DUMP_SHAPES=gpurun_out/r02_gemm_shapes_B98.json run 420 r02_bench_cfg3_N1.log python bench.py --steps 10 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg3_N1.log
run 300 r02_bench_torch_gpu.log python tools/bench_torch_gpu.py 128 2,14,98; tail -n 5 gpurun_out/r02_bench_torch_gpu.log
run 300 r02_bench_cfg2_N1.log python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline; tail -n 2 gpurun_out/r02_bench_cfg2_N1.log
run 240 r02_bench_cfg5_N1.log python bench.py --config cfg5 --steps 4 --warmup 4 --no-cpu-baseline; tail -n 2 gpurun_out/r02_bench_cfg5_N1.log
BENCH_BREAKDOWN=1 BENCH_B=98 run 240 r02_breakdown_batch98.log python tools/bench_denoiser.py 128; tail -n 32 gpurun_out/r02_breakdown_batch98.log
run 200 r02_bench_vae.log python tools/bench_vae.py 4096; tail -n 6 gpurun_out/r02_bench_vae.log
bash tools/gpu_profile_r02.sh
