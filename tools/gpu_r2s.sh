#!/bin/bash
# A/B: persistent tile loop of the P2G kernel (next tile's x loaded before the scatter walk)
set -u
mkdir -p gpurun_out
L=softmac_b200/lib
bash tools/gpu_variants.sh r2s "rest" $L/var_pers.so $L/var_pers.so,SMX_PERSIST=1 $L/var_pers.so $L/var_pers.so,SMX_PERSIST=1
bash tools/gpu_variants.sh r2s "stressed" $L/var_pers.so $L/var_pers.so,SMX_PERSIST=1
SMX_LIB=$PWD/$L/var_pers.so SMX_PERSIST=1 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "rollout or fusion or properties" 2>&1 | tail -5
