#!/bin/bash
# Round-2 GPU pass X (1 GPU): final state -- strict suite, default bench line, ncu launch list of the same command, smoke().
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -3 gpurun_out/x_pytest.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/x_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-standin --no-batch8 --no-config2 > gpurun_out/launches_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<PY
import json
d = json.loads(open("gpurun_out/x_bench.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "graph", (d["e2e"].get("graph_replay") or {}).get("value"), "clocks", d["clocks"])
print("train_step", {k: d["train_step"][k].get("ms_per_step") for k in ("fused", "fused_assembled_sh", "torch") if k in d["train_step"]}, d["train_step"].get("hbm_frac"))
print("standin", d["gpu_standin_baseline"]["ms_per_step"], "config2", json.dumps(d["config2"])[:500])
PY
