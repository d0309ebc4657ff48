#!/bin/bash
# A/B of several environment settings on the N=1 bench line: each argument is "VAR=val[,VAR2=val2]" (quote it); "base" = defaults
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  envs=""
  if [ "$spec" != "base" ]; then envs=$(echo "$spec" | tr ';' ' '); fi
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-e2e --no-config2 > gpurun_out/abm_$i.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/abm_$i.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("$spec", round(d["value"], 1), round(d["ms_per_step"], 4), "fwd", round(k["render_forward"]["ms_per_step"], 4), "bwd", round(k["render_backward"]["ms_per_step"], 4), "pro", round(k["backward_prologue"]["ms_per_step"], 4))
PY
done
