#!/bin/bash
# Round-2 GPU pass E (2 GPUs): row-sparse NVLS bench at N=2 (+ dense, nccl for comparison) and, on GPU 0, the async test
# + the N=1 bench line with the config2 entry.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench_$N.json 2> gpurun_out/e_bench_$N.err
echo "bench rc=$?"; grep -A6 "Traceback" gpurun_out/e_bench_$N.err | head -20
python - <<PY
import json
for tag in ("",):
    try:
        d = json.loads(open(f"gpurun_out/e_bench_$N{tag}.json").read().strip().splitlines()[-1])
        print(tag or "_sparse", d["value"], d["ms_per_step"], d["config"]["collective"], d.get("allreduce_check"), (d.get("batch8") or {}).get("views_s"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(tag, "failed", e)
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -x -k "async or binning or back_to_back" 2>&1 | tail -4
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step > gpurun_out/e_bench_1.json 2> gpurun_out/e_bench_1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/e_bench_1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "config2", d["config2"], "batch8", d["batch8"]["views_s"])
PY
