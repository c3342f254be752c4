#!/bin/bash
# round-2 GPU session 3b (1 GPU, short): attention tile-stagger sweep; epilogue-mode A/B on the whole batch-98 denoiser call.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 150 r02_selftest_attnstagger.log tools/selftest attnstagger; cat gpurun_out/r02_selftest_attnstagger.log
for mode in auto 0 1; do
  if [ "$mode" = auto ]; then unset SUPIR_B200_GEMM_WARP_EPILOGUE; else export SUPIR_B200_GEMM_WARP_EPILOGUE=$mode; fi
  BENCH_BREAKDOWN=1 BENCH_B=98 run 200 r02_breakdown_epi_$mode.log python tools/bench_denoiser.py 128
  grep -E "gemm  |conv3x3  |conv_geom|\"latent\"" gpurun_out/r02_breakdown_epi_$mode.log
done
unset SUPIR_B200_GEMM_WARP_EPILOGUE
SUPIR_BENCH_DUMP_SHAPES=gpurun_out/r02_gemm_shapes_B98.json run 300 r02_bench_cfg3_N1_b.log python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-run; tail -n 2 gpurun_out/r02_bench_cfg3_N1_b.log | cut -c1-600
