#!/bin/bash
set -u
mkdir -p gpurun_out
L=softmac_b200/lib
bash tools/gpu_variants.sh r2q "rest stressed" $L/var_cur.so $L/var_m.so $L/var_fast.so
bash tools/gpu_variants.sh r2q "rest" $L/var_cur.so $L/var_m.so
SMX_LIB=$PWD/$L/var_fastfull.so python -m pytest tests -m gpu -q 2>&1 | tail -8
