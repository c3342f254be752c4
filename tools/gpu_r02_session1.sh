#!/bin/bash
# round-2 GPU session 1: new persistent attention kernel (parity for the exponent-emulation settings, throughput sweep),
# the benchmarked-shape parity suite, the existing GPU suite, one short bench line. Every stage has its own timeout and log.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_s1_smi.txt 2>&1
stdbuf -oL timeout 300 tools/selftest attn > gpurun_out/r02_selftest_attn.log 2>&1; echo "exit=$?" >> gpurun_out/r02_selftest_attn.log
tail -n 45 gpurun_out/r02_selftest_attn.log
if grep -q "exit=0" gpurun_out/r02_selftest_attn.log; then
  stdbuf -oL timeout 300 tools/selftest attnperf2 > gpurun_out/r02_selftest_attnperf2.log 2>&1; echo "exit=$?" >> gpurun_out/r02_selftest_attnperf2.log
  cat gpurun_out/r02_selftest_attnperf2.log
fi
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -x -q -m gpu -s > gpurun_out/r02_pytest_bench_shapes.log 2>&1; echo "exit=$?" >> gpurun_out/r02_pytest_bench_shapes.log
tail -n 40 gpurun_out/r02_pytest_bench_shapes.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_bench_shapes.py --ignore=tests/test_gpu_bench_shapes.py -s --durations=15 > gpurun_out/r02_pytest_gpu_rest.log 2>&1; echo "exit=$?" >> gpurun_out/r02_pytest_gpu_rest.log
tail -n 40 gpurun_out/r02_pytest_gpu_rest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_s1.json 2> gpurun_out/r02_bench_s1.err; echo "bench exit=$?"
cat gpurun_out/r02_bench_s1.json; tail -n 5 gpurun_out/r02_bench_s1.err
