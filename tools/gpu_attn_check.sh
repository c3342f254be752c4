#!/bin/bash
# attention kernel check on the GPU box: unbuffered logs so that a hang still leaves evidence
mkdir -p gpurun_out
stdbuf -oL timeout 60 tools/selftest attn > gpurun_out/selftest_attn.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attn.log
tail -9 gpurun_out/selftest_attn.log
if grep -q "exit=0" gpurun_out/selftest_attn.log; then
  stdbuf -oL timeout 60 tools/selftest attnperf > gpurun_out/selftest_attnperf.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attnperf.log
  tail -7 gpurun_out/selftest_attnperf.log
  BENCH_B=14 timeout 150 python tools/bench_denoiser.py 128 2>&1 | tail -1
fi
