"""Runs the fused photometric loss forward + backward a few times at 1080p (profiling target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scgaussian_b200.losses import photometric_loss
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
y = torch.rand(3, 1080, 1920, generator=g).to(dev)
x = (y + 0.1 * torch.randn(3, 1080, 1920, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    loss = photometric_loss(x, y, 0.2)
    loss.backward()
    x.grad = None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    loss = photometric_loss(x, y, 0.2)
    loss.backward()
    x.grad = None
e1.record()
torch.cuda.synchronize()
print("fwd+bwd us:", e0.elapsed_time(e1) / 20 * 1000, float(loss))
