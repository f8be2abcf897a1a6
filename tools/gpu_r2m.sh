#!/bin/bash
# round-2 re-entry: full GPU suite + bench line of the current library, then A/B of CTA-size and start-stagger variants
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r2m_pytest.txt
python bench.py > gpurun_out/r2m_bench_line.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2m_bench_line.json
L=softmac_b200/lib
bash tools/gpu_variants.sh r2m "rest" $L/libsoftmac_b200.so $L/libsoftmac_b200.so,SMX_DBG=76800 $L/libsoftmac_b200.so,SMX_DBG=128000 $L/libsoftmac_b200.so,SMX_DBG=204800 \
   $L/var_w1.so $L/var_w1.so,SMX_DBG=25600 $L/var_w1.so,SMX_DBG=128000 $L/var_w2.so $L/var_w2.so,SMX_DBG=64000
