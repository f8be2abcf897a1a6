#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_coupling_episode.py -m gpu -q -k door_like 2>&1 | tail -30 | tee gpurun_out/r2n_door.txt
( time python -m pytest tests -m gpu -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r2n_pytest.txt
