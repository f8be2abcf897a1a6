#!/bin/bash
set -u
mkdir -p gpurun_out
STEPS=12 python tools/debug_door.py 2>&1 | grep weights | tee gpurun_out/r2o_door_debug.txt
python -m pytest tests/test_cuda_parity.py tests/test_coupling_episode.py -m gpu -q -k "von_mises or door_like" 2>&1 | tail -30 | tee gpurun_out/r2o_vm.txt
