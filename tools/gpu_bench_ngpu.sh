#!/bin/bash
# round-2: the driver's N-GPU launch of bench.py (headline arm + rollouts64 + slab_8M sub-records).  Usage: bash tools/gpu_bench_ngpu.sh N [extra bench args]
set -u
N=${1:-2}; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 "$@" \
  > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2e_bench_n$N.json"))
    print("value %.3f G/s  e2e %.3f (seq %.3f, f64 %.3f)" % (d["value"] / 1e9, d["e2e"]["value"] / 1e9, d["e2e"]["sequential_value"] / 1e9, d["e2e"]["f64"]["value"] / 1e9))
    for k in ("rollouts64", "slab_8M"):
        print(k, json.dumps(d.get(k))[:1800])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2e_bench_n$N.err").read()[-3000:])
PY
tail -5 gpurun_out/r2e_bench_n$N.err
