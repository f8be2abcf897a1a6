#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_cuda_batch.py -m gpu -q -k "graph" 2>&1 | tail -3
python tools/bench_demo.py --config grip --parity-env-steps 0 --batch 1 --sort-every 25 --arms device,device_graph > gpurun_out/r2_demo_grip_graph_b1.json 2> gpurun_out/r2k_grip.err; echo "grip rc=$?"
python tools/bench_demo.py --config pour --pour-actions adjusted --parity-env-steps 0 --arms device,device_graph > gpurun_out/r2_demo_pour_graph.json 2> gpurun_out/r2k_pour.err; echo "pour rc=$?"
python - <<'PY'
import json
for f in ("r2_demo_grip_graph_b1", "r2_demo_pour_graph"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        for a, r in d["arms"].items():
            print(f, a, "B=%d" % r["rollouts_in_handle"], "%.1f us/pair" % r["us_per_substep_pair"], "%.3g p-substeps/s" % r["particle_substeps_per_s_fwd_bwd"], "loss %.6g" % r["loss"], "|g| %.6g" % r["grad_norm"], r.get("graph", ""))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/r2k_{'grip' if 'grip' in f else 'pour'}.err").read()[-1500:])
PY
python tools/bench_slabs.py --substeps 16 2>&1 | grep "^{" | tail -1 | tee gpurun_out/r2l_slab_8M_n1.json | cut -c1-400
