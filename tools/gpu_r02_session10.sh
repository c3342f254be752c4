#!/bin/bash
# round-2 GPU session 10 (4 GPUs): the default workload at N=4 (own scaling table: 1 / 2 / 4 / 8).
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
run 300 r02_bench_cfg3_N4.log $TR bench.py --gpus 4 --steps 10 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg3_N4.log | cut -c1-900
