#!/bin/bash
# round-2 GPU session 7 (2 GPUs, short): bit-identity of the sharded paths after the tiled-VAE change, and the VAE time at N=2.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
run 240 r02_multigpu_check_N2_b.log $TR tools/check_multigpu.py; tail -n 2 gpurun_out/r02_multigpu_check_N2_b.log
run 300 r02_bench_cfg3_N2_b.log $TR bench.py --gpus 2 --steps 10 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg3_N2_b.log | cut -c1-1200
