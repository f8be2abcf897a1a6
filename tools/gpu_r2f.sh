#!/bin/bash
# round-2: peer-memory halo exchange: emulated-rank tests (1 GPU), then the 2-GPU NCCL-vs-peer comparison
set -u
mkdir -p gpurun_out
python -m pytest tests/test_cuda_slabs.py tests/test_cuda_parity.py tests/test_coupling_episode.py tests/test_cuda_batch.py -m gpu -x -q 2>&1 | tail -25
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
  for mode in "" "--nccl"; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/bench_slabs.py --check $mode 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/bench_slabs.py --substeps 16 $mode 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -2 | tee gpurun_out/r2f_slab_n${N}${mode}.json
  done
fi
