# GPU box: tests, default bench, ncu launch list and ncu --set full captures -> gpurun_out/
# (then, back in the build container: python tools/summarize_ncu.py rNN
#                                     python tools/summarize_ncu.py rNN_model gpurun_out/prof_model.ncu-rep)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^(backward_prologue|depth_key|emit_instances|init_ranges|onesweep_pass|preprocess_|render_|scan_offsets)" -s 60 -c 15 -o gpurun_out/prof_all -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_all.log 2>&1
# the model passes around the rasterizer (assembly, Adam) and the whole training step
python tools/model_step_time.py > gpurun_out/model_time.log 2>&1
python tools/train_step_time.py > gpurun_out/train_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_|adam_" -s 6 -c 3 -o gpurun_out/prof_model -f python tools/model_ncu.py > gpurun_out/prof_model.log 2>&1
cat gpurun_out/pytest_gpu.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"])
PY
ls -la gpurun_out | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
