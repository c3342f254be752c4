#!/bin/bash
# attention kernel check on the GPU box: unbuffered logs so that a hang still leaves evidence
mkdir -p gpurun_out
for single in 0 1; do
  export SUPIR_B200_ATTN_SINGLE=$single
  echo "== SUPIR_B200_ATTN_SINGLE=$single"
  stdbuf -oL timeout 60 tools/selftest attn > gpurun_out/selftest_attn_$single.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attn_$single.log
  tail -9 gpurun_out/selftest_attn_$single.log
  if grep -q "exit=0" gpurun_out/selftest_attn_$single.log; then
    stdbuf -oL timeout 60 tools/selftest attnperf2 > gpurun_out/selftest_attnperf2_$single.log 2>&1; echo "exit=$?" >> gpurun_out/selftest_attnperf2_$single.log
    tail -6 gpurun_out/selftest_attnperf2_$single.log
    BENCH_B=98 timeout 150 python tools/bench_denoiser.py 128 2>&1 | tail -1
  fi
done
