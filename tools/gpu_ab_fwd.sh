# A/B of the render_forward staging variants (SCGR_TMA_FWD: 0 register prefetch, 3 cp.async) x SCGR_FWD_MINB; backward on cp.async
for cfg in "0 20" "3 20" "3 22" "3 24" "3 26"; do set -- $cfg
SCGR_TMA_BWD=2 SCGR_TMA_FWD=$1 SCGR_FWD_MINB=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-e2e --no-config2 > gpurun_out/ab_fwd_$1_$2.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_fwd_$1_$2.json").read().strip().splitlines()[-1])
print("TMA_FWD=$1 MINB=$2", round(d["value"], 1), round(d["ms_per_step"], 4), "render_forward", round(d["kernels"]["render_forward"]["ms_per_step"], 4), "render_backward", round(d["kernels"]["render_backward"]["ms_per_step"], 4))
PY
done
