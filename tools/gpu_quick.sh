#!/bin/bash
# Quick 1-GPU check during kernel work: the structural / parity tests selected by $1 (pytest -k), then the N=1 bench line
# without its side entries; prints the per-kernel times.  TAG=$2 names the output files.
set -u
K=${1:-"config3 or forward_stages or binning"}
TAG=${2:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["samples"])
print("kernels", json.dumps({k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()}))
PY
