#!/bin/bash
# round-2: full GPU test suite + the driver's bench line (new e2e arm)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2d_pytest.log
python bench.py --steps 6 --warmup 3 --no-subrecords > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench.json"))
    print("value %.3f G/s" % (d["value"] / 1e9)); print(json.dumps(d["e2e"], indent=1)[:2500])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2d_bench.err").read()[-1500:])
PY
