#!/bin/bash
# round-2 GPU session 4 (2 GPUs): sharded == single-GPU bit for bit (tiled sampler units, untiled CFG-branch split, tiled VAE),
# then the bench lines at N=2.
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; stdbuf -oL timeout "$t" "$@" > "gpurun_out/$log" 2>&1; local rc=$?; echo "exit=$rc" >> "gpurun_out/$log"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run 300 r02_multigpu_check_N2.log $TR tools/check_multigpu.py; tail -n 4 gpurun_out/r02_multigpu_check_N2.log
run 360 r02_bench_cfg3_N2.log $TR bench.py --gpus 2 --steps 10 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg3_N2.log | cut -c1-1500
run 300 r02_bench_cfg2pair_N2.log $TR bench.py --gpus 2 --config cfg2pair --steps 5 --warmup 3; tail -n 2 gpurun_out/r02_bench_cfg2pair_N2.log | cut -c1-1500
run 420 r02_bench_cfg4_N2.log $TR bench.py --gpus 2 --config cfg4 --steps 2 --warmup 3 --no-full-run; tail -n 2 gpurun_out/r02_bench_cfg4_N2.log | cut -c1-1200
