#!/bin/bash
# A/B on one box: the library of commit 3e90fc3 (its own tree under _ab_3e90/) against the current one, and a fast-math build
set -u
mkdir -p gpurun_out
L=softmac_b200/lib
old() { (cd _ab_3e90 && python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-subrecords 2>/dev/null) > gpurun_out/r2p_old_$1.json; python -c "
import json; d=json.load(open('gpurun_out/r2p_old_$1.json')); k=d['roofline']['kernel_ms']
print('old($1): %.3f G/s %.2f ms ' % (d['value']/1e9, d['ms_per_step']) + ' '.join('%s=%.1f' % (a, 1e3*b) for a, b in sorted(k.items())))"; }
old a
bash tools/gpu_variants.sh r2p "rest" $L/libsoftmac_b200.so $L/var_cur.so $L/var_fast.so
old b
bash tools/gpu_variants.sh r2p "stressed" $L/var_cur.so $L/var_fast.so
SMX_LIB=$PWD/$L/var_fast.so python -m pytest tests/test_cuda_parity.py tests/test_cuda_branches.py -m gpu -q -x 2>&1 | tail -8
