#!/bin/bash
# round-2 GPU session 11 (1 GPU): ncu evidence from the FINAL tree (launch list of one timed step, full captures).
bash tools/gpu_profile_r02.sh
