#!/bin/bash
# round-2 GPU session 12 (1 GPU, the round's last GPU minutes): the tests added in this session — text conditioner kernels and
# towers, prompts -> image through the engine, fast-mode tiled VAE — plus the VAE / engine tests their host code touches, and smoke().
mkdir -p gpurun_out
( timeout 340 python -m pytest tests/test_gpu_conditioner.py tests/test_gpu_model.py tests/test_gpu_vae_sampler.py -q -s -m gpu 2>&1 | tail -150 ) > gpurun_out/r02_s12_new_tests.txt
( timeout 70 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/r02_s12_smoke.txt
tail -30 gpurun_out/r02_s12_new_tests.txt; cat gpurun_out/r02_s12_smoke.txt
