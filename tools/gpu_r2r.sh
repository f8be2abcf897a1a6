#!/bin/bash
# A/B: streaming loads that do not allocate in L1; L1 / shared split of the scatter kernels
set -u
mkdir -p gpurun_out
L=softmac_b200/lib
bash tools/gpu_variants.sh r2r "rest" $L/var_cur.so $L/var_noalloc.so $L/var_noalloc2.so $L/var_cur.so,SMX_CARVEOUT_FB=100 $L/var_cur.so,SMX_CARVEOUT_FB=75 $L/var_cur.so,SMX_CARVEOUT_FB=50 $L/var_sc3.so $L/var_sc3.so,SMX_CARVEOUT=75 $L/var_cur.so
bash tools/gpu_variants.sh r2r "stressed" $L/var_cur.so $L/var_noalloc2.so
