#!/bin/bash
# Round-2 GPU pass Q (N GPUs): fused NVLS collective (one launch, in-kernel barriers) -- tests, then bench fused / unfused.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nvls_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/q_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest_$N.log; tail -12 gpurun_out/q_pytest_$N.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for F in 1 0; do
SCGR_ALLREDUCE_FUSED=$F timeout 600 $TR --master-port 2951$F bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-batch8 > gpurun_out/q_bench_${N}_f$F.json 2> gpurun_out/q_bench_${N}_f$F.err
echo "bench fused=$F rc=$?"; grep -A8 "Traceback" gpurun_out/q_bench_${N}_f$F.err | head -20
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q_bench_${N}_f$F.json").read().strip().splitlines()[-1])
    print("fused=$F", d["value"], d["ms_per_step"], d["config"]["collective"], (d.get("allreduce_check") or {}).get("ok"))
except Exception as e:
    print("fused=$F failed", e)
PY
done
