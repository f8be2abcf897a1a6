#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/ubench_red | tee gpurun_out/r2i_ubench_red.txt
bash tools/gpu_variants.sh r2i "rest" softmac_b200/lib/libsoftmac_b200.so
python -m pytest tests/test_cuda_parity.py tests/test_cuda_slabs.py -m gpu -q 2>&1 | tail -4
