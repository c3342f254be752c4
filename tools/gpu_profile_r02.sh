#!/bin/bash
# ncu evidence for profiles/ (round 2): (1) launch list of ONE timed bench step (cudaProfilerStart/Stop around the timed
# region), (2) full-set captures of the dominant GEMM and of the self-attention kernel in situ at batch 98.
mkdir -p gpurun_out
export SUPIR_BENCH_SKIP_VAE=1 SUPIR_BENCH_CUDA_PROFILER=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --profile-from-start off \
  --csv --log-file gpurun_out/r02_launches_bench_step.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-run > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list exit=$?"; tail -c 400 gpurun_out/r02_ncu_bench.log
unset SUPIR_BENCH_SKIP_VAE SUPIR_BENCH_CUDA_PROFILER
python tools/summarize_launches.py gpurun_out/r02_launches_bench_step.csv > gpurun_out/r02_launches_bench_step.summary.txt 2>&1; head -30 gpurun_out/r02_launches_bench_step.summary.txt
BENCH_B=98 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 200 --launch-count 3 \
  -o gpurun_out/r02_prof_gemm -f python tools/bench_denoiser.py 128 > gpurun_out/r02_ncu_full_gemm.log 2>&1
echo "gemm full exit=$?"
BENCH_B=98 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_d64 --launch-skip 20 --launch-count 4 \
  -o gpurun_out/r02_prof_attn -f python tools/bench_denoiser.py 128 > gpurun_out/r02_ncu_full_attn.log 2>&1
echo "attn full exit=$?"
ls -la gpurun_out/*.ncu-rep
