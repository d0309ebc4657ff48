"""Times the fused photometric loss forward and backward at 1080p (and 4K) under both kernel designs
(SCGR_LOSS_VARIANT, read by the library on every launch) and a few band counts of the streaming one.
    python tools/loss_only.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scgaussian_b200.losses import photometric_loss
from scgaussian_b200 import _lib

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lib = _lib.load()


def run(shape, env):
    for k in ("SCGR_LOSS_VARIANT", "SCGR_LOSS_BANDS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = torch.Generator().manual_seed(0)
    y = torch.rand(*shape, generator=g).to(dev)
    x = (y + 0.1 * torch.randn(*shape, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
    for _ in range(3):
        photometric_loss(x, y, 0.2).backward()
        x.grad = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        ev[0].record()
        loss = photometric_loss(x, y, 0.2)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[1].elapsed_time(ev[2])
        x.grad = None
    print(f"{shape} {env}: forward {tf / reps * 1000:.1f} us, backward {tb / reps * 1000:.1f} us (events around the "
          f"public calls: includes torch's autograd bookkeeping), loss {float(loss):.7f}", flush=True)


for shape in ((3, 1080, 1920), (3, 2160, 3840)):
    run(shape, {"SCGR_LOSS_VARIANT": "0"})
    run(shape, {"SCGR_LOSS_VARIANT": "1"})
    for bands in (8, 14, 22):
        run(shape, {"SCGR_LOSS_VARIANT": "1", "SCGR_LOSS_BANDS": str(bands)})
