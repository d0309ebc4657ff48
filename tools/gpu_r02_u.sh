#!/bin/bash
# Round-2 GPU pass U (1 GPU): the loss kernels under ncu (durations of both designs, --set full of the streaming pair),
# loss timing tool, end-to-end entry with the new defaults.
set -u
mkdir -p gpurun_out
timeout 200 python tools/loss_only.py 20 > gpurun_out/u_loss_only.log 2>&1; grep -v Warning gpurun_out/u_loss_only.log | cut -c1-200
cat > /tmp/loss_ncu.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from scgaussian_b200.losses import photometric_loss
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
y = torch.rand(3, 1080, 1920, generator=g).to(dev)
x = (y + 0.1 * torch.randn(3, 1080, 1920, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
for v in ("0", "1", "0", "1", "0", "1"):
    os.environ["SCGR_LOSS_VARIANT"] = v
    photometric_loss(x, y, 0.2).backward(); x.grad = None
torch.cuda.synchronize()
PY
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:photometric --csv --log-file gpurun_out/u_loss_launches.csv python /tmp/loss_ncu.py > gpurun_out/u_loss_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:photometric.*stream -s 2 -c 2 -o gpurun_out/u_loss_stream -f python /tmp/loss_ncu.py >> gpurun_out/u_loss_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/u_loss_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iid = hdr.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((r[iid], r[ik][:60]), {})[r[im]] = r[iv]
for k, v in out.items():
    print(k, v)
PY
FAST="--no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-config2"
timeout 300 python bench.py --steps 24 --warmup 8 $FAST > gpurun_out/u_e2e.json 2> gpurun_out/u_e2e.err || tail -c 400 gpurun_out/u_e2e.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_e2e.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "graph", (d["e2e"].get("graph_replay") or {}).get("value"))
PY
