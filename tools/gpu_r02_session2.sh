#!/bin/bash
# round-2 GPU session 2 (fail-fast): hang guard first — if the quick attention check does not pass inside 90 s nothing else
# runs. Then attention parity + throughput sweep, epilogue-mode sweep, benchmarked-shape parity, the rest of the GPU suite,
# one bench line. Every stage has its own timeout and log; pytest additionally has a per-test timeout.
mkdir -p gpurun_out
run() {  # run <timeout_s> <logfile> <command...>
  local t=$1 log=$2; shift 2
  stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?
  echo "exit=$rc" >> "gpurun_out/$log"
  return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_s2_smi.txt 2>&1
run 90 r02_selftest_attnquick.log tools/selftest attnquick || { tail -n 30 gpurun_out/r02_selftest_attnquick.log; echo "HANG GUARD FAILED - stopping"; exit 1; }
tail -n 8 gpurun_out/r02_selftest_attnquick.log
# geometry-mode convolutions (TMA element strides): if this does not pass, the rest of the session uses the im2col / upsample path
if run 150 r02_pytest_conv_geom.log python -m pytest tests/test_gpu_ops.py -q -m gpu -k "strided_and_subpixel" -s --timeout 100; then echo "conv geometry: ok"; else
  tail -n 20 gpurun_out/r02_pytest_conv_geom.log; echo "conv geometry FAILED: continuing with SUPIR_B200_CONV_GEOM=0"; export SUPIR_B200_CONV_GEOM=0; fi
run 200 r02_selftest_attn.log tools/selftest attn; tail -n 12 gpurun_out/r02_selftest_attn.log
run 200 r02_selftest_attnperf2.log tools/selftest attnperf2; cat gpurun_out/r02_selftest_attnperf2.log
run 200 r02_selftest_epiperf.log tools/selftest epiperf; cat gpurun_out/r02_selftest_epiperf.log
run 500 r02_pytest_bench_shapes.log python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -s --timeout 150; tail -n 30 gpurun_out/r02_pytest_bench_shapes.log
run 420 r02_pytest_gpu_ops.log python -m pytest tests -q -m gpu --ignore=tests/test_gpu_bench_shapes.py --ignore=tests/test_gpu_bench_networks.py -s --timeout 150 --durations=10; tail -n 40 gpurun_out/r02_pytest_gpu_ops.log
run 600 r02_pytest_networks.log python -m pytest tests/test_gpu_bench_networks.py -q -m gpu -s --timeout 420 --durations=10; tail -n 30 gpurun_out/r02_pytest_networks.log
run 360 r02_bench_s2.log python bench.py --steps 5 --warmup 3 --no-cpu-baseline; tail -n 3 gpurun_out/r02_bench_s2.log
