#!/bin/bash
# Round-2 GPU pass H (1 GPU): strict suite (split SH layout tests included), bench (clock sampler across warm-up + timed
# region; train_step with the split / assembled SH variants), ncu of the preprocess kernels in the split train step.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -6 gpurun_out/h_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/h_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/h_bench.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
ts = d["train_step"]
print("train_step", {k: ts[k].get("ms_per_step") for k in ("fused", "fused_assembled_sh", "torch") if k in ts}, ts.get("hbm_frac"), ts.get("hbm_frac_assembled_sh"))
for k in ("fused", "fused_assembled_sh"):
    print(k, json.dumps(ts.get(k, {}).get("kernel_us_per_step")))
print("kernels", json.dumps({k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()}))
PY
VARIANTS=fused STEPS=4 WARM=3 timeout 900 ncu --set full --clock-control none -k regex:"preprocess_forward|preprocess_backward|assemble_" -s 12 -c 8 -f -o gpurun_out/h_split python tools/train_step_time.py > gpurun_out/h_ncu_split.log 2>&1
ls -la gpurun_out/h_split.ncu-rep
