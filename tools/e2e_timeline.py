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
from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402
from scgaussian_b200 import rasterizer as R  # noqa: E402
from scgaussian_b200.losses import photometric_loss  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "e2e"
dev = torch.device("cuda", 0)
cam, sc, grads = B.make_inputs(0, dev)
t = {k: v.to(dev).contiguous() for k, v in sc.items()}
gC, gD, gA = [g.to(dev).contiguous() for g in grads]
H, W, P = B.HEIGHT, B.WIDTH, B.P_GAUSS
s = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], torch.zeros(3, device=dev), 1.0,
                                  cam["viewmatrix"].to(dev), cam["projmatrix"].to(dev), B.SH_DEG,
                                  cam["campos"].to(dev), False, False)
gt_host = torch.rand(3, H, W).pin_memory()
vm_host, pm_host = cam["viewmatrix"].clone().pin_memory(), cam["projmatrix"].clone().pin_memory()
cp_host, bg_host = cam["campos"].clone().pin_memory(), torch.zeros(3).pin_memory()
leaves = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
loss_host = torch.zeros(1).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
gt_dev = torch.empty(3, H, W, device=dev)
m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
args_in = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None)


def e2e_step():
    copy_stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(copy_stream):
        gt_dev.copy_(gt_host, non_blocking=True)
    s2 = s._replace(viewmatrix=vm_host.to(dev, non_blocking=True), projmatrix=pm_host.to(dev, non_blocking=True),
                    campos=cp_host.to(dev, non_blocking=True), bg=bg_host.to(dev, non_blocking=True))
    color, radii, depth, alpha = GaussianRasterizer(s2)(
        means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"])
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    loss = photometric_loss(color, gt_dev, 0.2) + (depth.sum() + alpha.sum()) * (0.01 / (H * W))
    loss.backward()
    m2d.grad = None
    loss_host.copy_(loss.detach().reshape(1), non_blocking=False)
    for v in leaves.values():
        v.grad = None


def raw_step():
    color, radii, depth, alpha, state = R.rasterize_forward_raw(*args_in, s)
    R.rasterize_backward_raw(state, *args_in, s, gC, gD, gA)


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
# last step only: find its first event = the (NS-1)-th occurrence of the first kernel name pattern
names = [e.name for e in ev]
key = "preprocess_forward"
idxs = [i for i, n in enumerate(names) if key in n]
start = idxs[-1]
# include the memcpys just before it
while start > 0 and ev[start - 1].time_range.start > ev[idxs[-2]].time_range.end + 0 and "preprocess_backward" not in ev[start - 1].name \
        and ev[idxs[-1]].time_range.start - ev[start - 1].time_range.start < 300:
    start -= 1
prev_end = ev[start].time_range.start
base = prev_end
for e in ev[start:]:
    gap = e.time_range.start - prev_end
    print(f"{e.time_range.start - base:9.1f} us  dur {e.time_range.end - e.time_range.start:8.1f}  gap {gap:7.1f}  {e.name[:90]}")
    prev_end = max(prev_end, e.time_range.end)
