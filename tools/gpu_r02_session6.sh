#!/bin/bash
# round-2 GPU session 6 (1 GPU): final validation the way the driver runs it — hang guard, ALU-pack sweep, the whole -m gpu suite,
# smoke(), the default bench line with the driver's flags, the reference arm.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
run 90 r02_selftest_attnquick6.log tools/selftest attnquick || { tail -n 20 gpurun_out/r02_selftest_attnquick6.log; echo "HANG GUARD FAILED"; exit 1; }
run 200 r02_selftest_attnpack.log tools/selftest attnpack; cat gpurun_out/r02_selftest_attnpack.log
run 900 r02_pytest_gpu_all.log python -m pytest tests -q -m gpu --timeout 420 --durations=12; tail -n 25 gpurun_out/r02_pytest_gpu_all.log
run 200 r02_smoke.log python -c "import __graft_entry__ as g; g.smoke()"; tail -n 3 gpurun_out/r02_smoke.log
run 420 r02_bench_final_N1.log python bench.py --gpus 1 --steps 20 --warmup 5; tail -n 1 gpurun_out/r02_bench_final_N1.log | cut -c1-1500
run 420 r02_bench_reference_arm.log python bench.py --impl reference --gpus 1 --steps 20 --warmup 5; tail -n 2 gpurun_out/r02_bench_reference_arm.log | cut -c1-1500
