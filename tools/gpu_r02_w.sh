#!/bin/bash
# Round-2 GPU pass W (N GPUs, short): programmatic dependent launch next to the fused collective -- NVLS tests, then the
# bench line with its allreduce_check (no extras).
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nvls_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/w_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w_pytest_$N.log; tail -3 gpurun_out/w_pytest_$N.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --no-standin --no-config2 > gpurun_out/w_bench_$N.json 2> gpurun_out/w_bench_$N.err
echo "bench rc=$?"; grep -A8 "Traceback" gpurun_out/w_bench_$N.err | head -24
python - <<PY
import json
d = json.loads(open("gpurun_out/w_bench_$N.json").read().strip().splitlines()[-1])
print("N=$N", d["value"], d["ms_per_step"], d.get("allreduce_check"), "batch8", (d.get("batch8") or {}).get("views_s"), "e2e", (d.get("e2e") or {}).get("value"))
PY
