#!/bin/bash
# round-2 GPU session 13 (1 GPU, the last GPU minutes of the round): engine (c / uc given and prompts -> image through the text
# conditioner) + VAE / sampler / Brownian-noise tests on the committed tree, then the conditioner timing at its real size.
mkdir -p gpurun_out
( timeout 170 python -m pytest tests/test_gpu_model.py tests/test_gpu_vae_sampler.py -q -s -m gpu 2>&1 | tail -60 ) > gpurun_out/r02_s13_engine_sampler_tests.txt
( timeout 80 python tools/bench_conditioner.py 2>&1 | tail -4 ) > gpurun_out/r02_s13_bench_conditioner.txt
tail -12 gpurun_out/r02_s13_engine_sampler_tests.txt; cat gpurun_out/r02_s13_bench_conditioner.txt
