#!/bin/bash
# Round-2 GPU pass P (1 GPU): the artefacts the tracked profiles are written from -- strict suite, default bench line,
# ncu launch list of the same command, ncu --set full of every rasterizer kernel of one step, smoke().
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/p_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-standin --no-batch8 --no-config2 > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(backward_prologue|depth_key|emit_instances|init_ranges|onesweep_pass|preprocess_|render_|scan_offsets)" -s 60 -c 15 -o gpurun_out/prof_all -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-standin --no-batch8 --no-config2 > gpurun_out/prof_all.log 2>&1
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:"render_" -s 6 -c 26 --csv --log-file gpurun_out/p_issue.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-standin --no-batch8 --no-config2 > gpurun_out/p_issue.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<PY
import json
d = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("train_step", {k: d["train_step"][k].get("ms_per_step") for k in ("fused", "fused_assembled_sh", "torch") if k in d["train_step"]}, d["train_step"].get("hbm_frac"))
print("standin", d["gpu_standin_baseline"]["ms_per_step"], "config2", json.dumps(d["config2"])[:500])
PY
ls -la gpurun_out/prof_all.ncu-rep gpurun_out/launches.csv gpurun_out/p_issue.csv
