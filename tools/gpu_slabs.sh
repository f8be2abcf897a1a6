#!/bin/bash
# round-2: slab decomposition on all GPUs of the box: correctness check + 8M / 256^3 strong scaling, peer memory vs the NCCL transport
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for mode in "" "--nccl"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/bench_slabs.py --check $mode 2>&1 | grep "^{" | tail -1 | tee gpurun_out/r2g_slab_check_n${N}${mode}.json | cut -c1-700
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/bench_slabs.py --substeps 16 $mode 2>&1 | grep "^{" | tail -1 | tee gpurun_out/r2g_slab_8M_n${N}${mode}.json | cut -c1-900
done
