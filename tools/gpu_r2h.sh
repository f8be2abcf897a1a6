#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2h_pytest.log
bash tools/gpu_variants.sh r2h "rest stressed" softmac_b200/lib/var_base.so softmac_b200/lib/libsoftmac_b200.so
