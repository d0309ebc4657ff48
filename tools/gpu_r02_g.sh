#!/bin/bash
# Round-2 GPU pass G (N GPUs): NVLS tests at N ranks, bench.py --gpus N (row-sparse NVLS; config4 at N = 8), NCCL run for comparison.
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nvls_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/g_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest_$N.log; tail -3 gpurun_out/g_pytest_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_bench_$N.json 2> gpurun_out/g_bench_$N.err
echo "bench rc=$?"; grep -A8 "Traceback" gpurun_out/g_bench_$N.err | head -24
SCGR_ALLREDUCE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-batch8 > gpurun_out/g_bench_${N}_nccl.json 2> gpurun_out/g_bench_${N}_nccl.err
python - <<PY
import json
for tag in ("", "_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/g_bench_$N{tag}.json").read().strip().splitlines()[-1])
        print(tag or "_sparse", d["value"], d["ms_per_step"], d["config"]["collective"], d.get("allreduce_check", {}).get("ok"), "batch8", (d.get("batch8") or {}).get("views_s"), "e2e", (d.get("e2e") or {}).get("value"))
        print("   config4", json.dumps(d.get("config4"))[:700])
    except Exception as e:
        print(tag, "failed", e)
PY
