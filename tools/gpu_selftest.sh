#!/bin/bash
# Runs the standalone kernel self-test on the GPU box; each group has its own timeout so a hung kernel cannot
# wedge the whole call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/selftest_smi.txt 2>&1
for g in "$@"; do
  echo "=== group $g" 
  timeout 240 tools/selftest $g > gpurun_out/selftest_$g.log 2>&1
  echo "exit=$?" >> gpurun_out/selftest_$g.log
  tail -n 60 gpurun_out/selftest_$g.log
done
