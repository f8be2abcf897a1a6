#!/bin/bash
# round-2: demo episodes (BASELINE configs 1 and 2) with the final library: timing arms + config 2 with the demo's own adjusted actions
set -u
mkdir -p gpurun_out
python tools/bench_demo.py --config grip --parity-env-steps 0 --batch 8 --sort-every 25 > gpurun_out/r2_demo_grip.json 2> gpurun_out/r2j_grip.err; echo "grip rc=$?"
python tools/bench_demo.py --config pour --pour-actions adjusted --parity-env-steps 300 --parity-strength 1.0 > gpurun_out/r2_demo_pour_adjusted.json 2> gpurun_out/r2j_pour.err; echo "pour rc=$?"
python - <<'PY'
import json
for f in ("r2_demo_grip", "r2_demo_pour_adjusted"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        for a, r in d["arms"].items():
            print(f, a, "B=%d" % r["rollouts_in_handle"], "%.1f us/pair" % r["us_per_substep_pair"], "%.3g p-substeps/s" % r["particle_substeps_per_s_fwd_bwd"], "grad finite", r["grad_finite"], "loss %.4g" % r["loss"])
        if "parity" in d: print(f, json.dumps(d["parity"]))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/r2j_{'grip' if 'grip' in f else 'pour'}.err").read()[-1500:])
PY
python -m pytest tests/test_cuda_parity.py -m gpu -q -k fp32_host 2>&1 | tail -2
