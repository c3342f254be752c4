#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one timed bench step, (2) full-set captures of the dominant GEMM and of attention
mkdir -p gpurun_out
export SUPIR_BENCH_SKIP_VAE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:supir \
  --launch-skip 3090 --launch-count 1600 --csv --log-file gpurun_out/launches_bench_step_r01b.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r01b.log 2>&1
echo "launch list exit=$?"; tail -c 600 gpurun_out/ncu_bench_r01b.log
unset SUPIR_BENCH_SKIP_VAE
BENCH_B=98 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 200 --launch-count 3 \
  -o gpurun_out/prof_gemm_r01b -f python tools/bench_denoiser.py 128 > gpurun_out/ncu_full_gemm_r01b.log 2>&1
echo "gemm full exit=$?"
BENCH_B=98 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_d64 --launch-skip 20 --launch-count 3 \
  -o gpurun_out/prof_attn_r01b -f python tools/bench_denoiser.py 128 > gpurun_out/ncu_full_attn_r01b.log 2>&1
echo "attn full exit=$?"
ls -la gpurun_out/*.ncu-rep
