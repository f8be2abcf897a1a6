#!/bin/bash
# round-2 measurement pass (run under gpurun): GPU tests, the driver's bench line, ncu launch list of the same command,
# ncu --set full of every substep kernel (+ the L2 reduction counters).  Usage: bash tools/gpu_measure.sh TAG [skiptests]
set -u
tag=${1:-r2}
mkdir -p gpurun_out
if [ "${2:-}" != "skiptests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
  tail -3 gpurun_out/${tag}_pytest.log
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/${tag}_bench.json
CMD="python bench.py --steps 1 --warmup 1 --substeps 16 --no-e2e --no-cpu-baseline --no-parity --no-subrecords"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --substeps 4 --no-e2e --no-cpu-baseline --no-parity --no-subrecords"
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum.per_second,lts__t_sectors_op_red.sum.per_second \
    --clock-control none --import-source on -k regex:'k_p2g|k_g2p|k_grid|k_bwd' -s 14 -c 14 -o gpurun_out/${tag}_prof $CMD2 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
# per-source-line executed instructions and stall samples of the two hot kernels (read with tools/ncu_src_summary.py)
for k in k_p2g_grad_g2p_grad k_p2g; do
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv -k regex:"^${k}" -c 1 > gpurun_out/${tag}_src_${k}.csv 2>/dev/null
done
ls -la gpurun_out | tail -6
