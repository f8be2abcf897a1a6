#!/bin/bash
# round-2 baseline: GPU tests, bench line, ncu --set full of the substep kernels (run under gpurun)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python bench.py --steps 4 --warmup 3 --init stressed --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_stressed.json 2> gpurun_out/r2a_bench_stressed.err; echo "bench stressed rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --substeps 4 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r2a_plain.log 2>&1 && \
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum \
    --clock-control none --import-source on -k regex:'k_p2g|k_g2p|k_grid' -s 21 -c 21 -o gpurun_out/r2a_prof $CMD > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench.json | cut -c1-600
