#!/bin/bash
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_bench.json"))
print("A %.3f stressed %.3f B %.3f e2e %.3f" % (d["value"] / 1e9, d["stressed"]["value"] / 1e9, d["variant_B"]["value"] / 1e9, d["e2e"]["value"] / 1e9), d["variant_B"]["parity"])
PY
