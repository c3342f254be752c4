#!/bin/bash
# attention kernel check on the GPU box (parity, then throughput at the step's batch, then one denoiser call).
# Logs are written unbuffered under gpurun_out/ so that a hang still leaves evidence; every stage has its own timeout.
mkdir -p gpurun_out
stdbuf -oL timeout 60 tools/selftest attn > gpurun_out/selftest_attn.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attn.log
tail -9 gpurun_out/selftest_attn.log
if grep -q "exit=0" gpurun_out/selftest_attn.log; then
  stdbuf -oL timeout 60 tools/selftest attnperf2 > gpurun_out/selftest_attnperf2.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attnperf2.log
  tail -6 gpurun_out/selftest_attnperf2.log
  BENCH_B=98 timeout 150 python tools/bench_denoiser.py 128 2>&1 | tail -1
fi
