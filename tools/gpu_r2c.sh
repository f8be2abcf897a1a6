#!/bin/bash
# round-2: new branch-coverage GPU tests + FULL-LENGTH episode parity of BASELINE configs 1 and 2 (oracle vs CUDA, same env loop)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_cuda_branches.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2c_pytest.log
python tools/bench_demo.py --config grip --arms "" --parity-env-steps 400 --parity-strength 1.0 > gpurun_out/r2_demo_grip_parity_full.json 2> gpurun_out/r2c_grip.err; echo "grip rc=$?"
python tools/bench_demo.py --config pour --arms "" --parity-env-steps 3000 --parity-strength 1.0 > gpurun_out/r2_demo_pour_parity_full.json 2> gpurun_out/r2c_pour.err; echo "pour rc=$?"
python - <<'PY'
import json
for c in ("grip", "pour"):
    try:
        d = json.load(open(f"gpurun_out/r2_demo_{c}_parity_full.json"))
        print(c, json.dumps(d["parity"]))
    except Exception as e:
        print(c, "FAILED", e); print(open(f"gpurun_out/r2c_{c}.err").read()[-800:])
PY
