#!/bin/bash
# A/B timing of library variants (built with `python softmac_b200/build.py --out softmac_b200/lib/var_X.so -D...`): one short
# device-resident bench line per variant.  Usage (under gpurun): bash tools/gpu_variants.sh tag "inits" spec ...
# spec = lib.so[,ENV=VAL[,ENV=VAL...]]   (e.g. softmac_b200/lib/libsoftmac_b200.so,SMX_DBG=2 : timing-only ablation, results invalid)
tag=$1; inits=$2; shift 2
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  lib=${parts[0]}; envs=("${parts[@]:1}")
  name=$(basename $lib .so)$(printf '_%s' "${envs[@]}" | tr -d '=' | sed 's/^_$//')
  for init in $inits; do
    out=gpurun_out/${tag}_${name}_${init}
    env SMX_LIB=$PWD/$lib "${envs[@]}" python bench.py --steps 6 --warmup 3 --init $init --no-e2e --no-cpu-baseline --no-parity --no-subrecords > $out.json 2> $out.err
    python - <<PY
import json
try:
    d = json.load(open("$out.json"))
    k = d["roofline"]["kernel_ms"]
    print("${name} ${init}: %.3f G/s  %.2f ms  " % (d["value"] / 1e9, d["ms_per_step"]) + " ".join("%s=%.1f" % (a, 1e3 * b) for a, b in sorted(k.items())))
except Exception as e:
    print("${name} ${init}: FAILED", e); print(open("$out.err").read()[-600:])
PY
  done
done
