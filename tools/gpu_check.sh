# GPU box: parity tests + default bench + serial-stage-1 A/B (run as: gpurun -- bash tools/gpu_check.sh)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["num_rendered_R"])
print({k:round(v["ms_per_step"]*1000,1) for k,v in d["kernels"].items()})
PY
cat gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_c.err
SCGR_SERIAL_STAGE1=1 python tools/ab.py "" 
