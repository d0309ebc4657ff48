"""GPU timeline of one end-to-end step (GPU box): runs the same step as bench.py's e2e leg under
torch.profiler and prints every kernel / memcpy of the last profiled step with its start offset,
duration and the idle gap before it.   python tools/e2e_timeline.py [raw|e2e]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench as B  # noqa: E402
from scgaussian_b200 import GaussianRasterizer  # noqa: E402
from scgaussian_b200.losses import photometric_loss  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "e2e"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = B.CONFIGS[3]
wl = B.Workload(cfg, dev, 0, 1)
H, W, P = cfg["H"], cfg["W"], cfg["P"]
gt_host = torch.rand(3, H, W).pin_memory()
cam_host = [torch.cat([c["viewmatrix"].reshape(-1), c["projmatrix"].reshape(-1), c["campos"].reshape(-1),
                       torch.zeros(3)]).contiguous().pin_memory() for c in wl.cams_cpu]
leaves = {k: wl.t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
loss_host = torch.zeros(1).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
gt_dev = torch.empty(3, H, W, device=dev)
cam_dev = torch.empty(38, device=dev)
m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
s0 = wl.settings[0]
counter = [0]


def e2e_step():
    k = counter[0]
    counter[0] += 1
    copy_stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(copy_stream):
        gt_dev.copy_(gt_host, non_blocking=True)
        cam_dev.copy_(cam_host[k % B.N_CAMERAS], non_blocking=True)
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    c = cam_dev
    s2 = s0._replace(viewmatrix=c[0:16].view(4, 4), projmatrix=c[16:32].view(4, 4), campos=c[32:35], bg=c[35:38])
    color, radii, depth, alpha = GaussianRasterizer(s2)(
        means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"])
    loss = photometric_loss(color, gt_dev, 0.2) + (depth.sum() + alpha.sum()) * (0.01 / (H * W))
    loss.backward()
    m2d.grad = None
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    for v in leaves.values():
        v.grad = None


def raw_step():
    k = counter[0]
    counter[0] += 1
    wl.view(k)


step = e2e_step if mode == "e2e" else raw_step
for _ in range(5):
    step()
torch.cuda.synchronize()
NS = 4
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tend = max(e.time_range.end for e in ev)
print(f"mode={mode}: {NS} steps span {(tend - t0) / NS:.1f} us/step on the GPU timeline; busy "
      f"{sum(e.time_range.end - e.time_range.start for e in ev) / NS:.1f} us/step (sum over streams)")
# last step only: from the last preprocess_forward (and the copies just before it) to the end
names = [e.name for e in ev]
idxs = [i for i, n in enumerate(names) if "preprocess_forward" in n]
start = idxs[-1]
while start > 0 and "preprocess_backward" not in ev[start - 1].name and ev[idxs[-1]].time_range.start - ev[start - 1].time_range.start < 300:
    start -= 1
prev_end = ev[start].time_range.start
base = prev_end
ours = other = 0.0
for e in ev[start:]:
    gap = e.time_range.start - prev_end
    dur = e.time_range.end - e.time_range.start
    if "scgr" in e.name:
        ours += dur
    else:
        other += dur
    print(f"{e.time_range.start - base:9.1f} us  dur {dur:8.1f}  gap {gap:7.1f}  {e.name[:90]}")
    prev_end = max(prev_end, e.time_range.end)
print(f"last step: libscgr kernels {ours:.1f} us, everything else (torch kernels, copies) {other:.1f} us, span {prev_end - base:.1f} us")
