#!/bin/bash
# A/B of an environment switch on the N=1 bench line (1 GPU): tools/gpu_ab_env.sh VAR v1 v2 ...   (TAG names the outputs)
set -u
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-e2e --no-config2 > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err || tail -c 600 gpurun_out/ab_${VAR}_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${VAR}_$v.json").read().strip().splitlines()[-1])
print("$VAR=$v", round(d["value"], 1), round(d["ms_per_step"], 4), json.dumps({k: round(x["ms_per_step"], 4) for k, x in d["kernels"].items()}))
PY
done
