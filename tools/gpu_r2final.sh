#!/bin/bash
# parity report of the final library + demo episodes with the final library (1 GPU)
set -u
mkdir -p gpurun_out
python tools/parity_report.py > gpurun_out/r2z_parity_report.txt 2>&1; tail -12 gpurun_out/r2z_parity_report.txt
