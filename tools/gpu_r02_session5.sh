#!/bin/bash
# round-2 GPU session 5 (8 GPUs): scaling of the default workload and BASELINE configs[3] (225 windows, 256 + 256 VAE tiles).
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
run 240 r02_bench_cfg3_N8.log $TR bench.py --gpus 8 --steps 10 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg3_N8.log
run 420 r02_bench_cfg4_N8.log $TR bench.py --gpus 8 --config cfg4 --steps 3 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg4_N8.log
