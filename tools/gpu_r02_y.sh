#!/bin/bash
# Round-2 GPU pass Y (1 GPU, < 1 min): the streaming loss kernels under ncu in their final form (durations + --set full).
set -u
mkdir -p gpurun_out
cat > /tmp/loss_ncu.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from scgaussian_b200.losses import photometric_loss
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
y = torch.rand(3, 1080, 1920, generator=g).to(dev)
x = (y + 0.1 * torch.randn(3, 1080, 1920, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
for _ in range(4):
    photometric_loss(x, y, 0.2).backward(); x.grad = None
torch.cuda.synchronize()
PY
timeout 60 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:photometric --csv --log-file gpurun_out/y_loss_launches.csv python /tmp/loss_ncu.py > gpurun_out/y_loss_ncu.log 2>&1
timeout 70 ncu --set full --clock-control none --import-source on -k regex:photometric.*stream -s 2 -c 2 -o gpurun_out/y_loss_stream -f python /tmp/loss_ncu.py >> gpurun_out/y_loss_ncu.log 2>&1
grep -E "photometric" gpurun_out/y_loss_launches.csv | cut -d, -f1,5,13- | tail -24
