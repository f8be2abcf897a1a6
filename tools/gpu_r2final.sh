#!/bin/bash
# timing experiment (results invalid): which access of k_grid_grad costs what (variant A)
set -u
mkdir -p gpurun_out
L=$PWD/softmac_b200/lib/var_dbg.so
for D in 0 8 16 32 56; do
  SMX_LIB=$L SMX_DBG=$D python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-subrecords > gpurun_out/r2y_dbg_$D.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/r2y_dbg_$D.json")); k = d["roofline"]["kernel_ms"]
print("dbg $D: %.3f G/s %.2f ms " % (d["value"] / 1e9, d["ms_per_step"]) + " ".join("%s=%.1f" % (a, 1e3 * k[a]) for a in ("k_grid_op", "k_grid_grad", "k_g2p2g", "k_p2g_grad+g2p_grad") if a in k))
PY
done
