#!/bin/bash
# Round-2 GPU pass F (1 GPU): strict suite, bench (config2 graph entry), ncu of the model / loss kernels in the train step.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/f_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "config2", json.dumps(d["config2"])[:900])
print("train_step", d["train_step"].get("fused", {}).get("ms_per_step"), d["train_step"].get("hbm_frac"))
PY
VARIANTS=fused STEPS=6 WARM=3 timeout 900 ncu --set full --clock-control none -k regex:"adam_kernel|assemble_|photometric_|densification_stats" -s 26 -c 12 -f -o gpurun_out/f_model python tools/train_step_time.py > gpurun_out/f_ncu_model.log 2>&1
ls -la gpurun_out/f_model.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
