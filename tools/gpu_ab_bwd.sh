for cfg in "1 18" "2 18" "2 20" "2 16"; do set -- $cfg
SCGR_TMA_BWD=$1 SCGR_BWD_MINB=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-e2e --no-config2 > gpurun_out/ab_bwd_$1_$2.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_bwd_$1_$2.json").read().strip().splitlines()[-1])
print("TMA_BWD=$1 MINB=$2", round(d["value"], 1), round(d["ms_per_step"], 4), "render_backward", round(d["kernels"]["render_backward"]["ms_per_step"], 4), "prologue", round(d["kernels"]["backward_prologue"]["ms_per_step"], 4))
PY
done
