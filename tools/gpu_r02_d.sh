#!/bin/bash
# Round-2 GPU pass D (N GPUs of one box): the NVLS collective tests, then bench.py under torchrun with the row-sparse
# NVLS all-reduce (default), the dense NVLS all-reduce and NCCL.   usage: bash tools/gpu_r02_d.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
timeout 900 python -m pytest tests/test_nvls_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/d_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest_$N.log
tail -6 gpurun_out/d_pytest_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/d_bench_$N.json 2> gpurun_out/d_bench_$N.err
echo "bench rc=$?"; tail -c 600 gpurun_out/d_bench_$N.err
X="--no-e2e --no-batch8"
SCGR_ALLREDUCE_SPARSE=0 timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 $X > gpurun_out/d_bench_${N}_dense.json 2> gpurun_out/d_bench_${N}_dense.err
SCGR_ALLREDUCE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 $X > gpurun_out/d_bench_${N}_nccl.json 2> gpurun_out/d_bench_${N}_nccl.err
python - <<PY
import json
for tag in ("", "_dense", "_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/d_bench_$N{tag}.json").read().strip().splitlines()[-1])
        print(tag or "_sparse", d["value"], d["ms_per_step"], d["config"]["collective"], d.get("allreduce_check"), (d.get("batch8") or {}).get("views_s"), (d.get("e2e") or {}).get("value"), (d.get("config4") or {}))
    except Exception as e:
        print(tag, "failed", e)
PY
