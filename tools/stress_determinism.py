"""Stress test (GPU box): the forward's binning must be bit-deterministic.  Runs the config-3 forward
N times and checks, every iteration: num_rendered, that the depth order is a permutation, that the
sum of tiles_touched equals num_rendered, and that the depth order / instance list / image equal
the first iteration's.  usage: python tools/stress_determinism.py [iters] [P] [W] [H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from scgaussian_b200 import synthetic as O  # noqa: E402
from scgaussian_b200 import GaussianRasterizationSettings  # noqa: E402
from scgaussian_b200 import rasterizer as R  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
dev = torch.device("cuda:0")
cam = O.make_camera(W, H)
sc = O.synth_scene(P, W, H, sh_degree=3, scale_median=0.01 if P >= 500_000 else 0.03, seed=0)
t = {k: v.to(dev).contiguous() for k, v in sc.items()}
s = GaussianRasterizationSettings(
    image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
    scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev),
    sh_degree=3, campos=cam["campos"].to(dev), prefiltered=False, debug=False)
args = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None)

ref = None
bad = 0
for it in range(iters):
    color, radii, depth, alpha, state = R.rasterize_forward_raw(*args, s)
    torch.cuda.synchronize()
    dv = R.debug_views(state, P, s)
    Rn = state.num_rendered
    order = dv["depth_order"].long()
    tt = dv["tiles_touched"].long()
    nvis = int((tt > 0).sum())
    cur = dict(R=Rn, sum_tt=int(tt.sum()), order=order.clone(), pl=dv["point_list"].clone(), color=color.clone(),
               ranges=dv["ranges"].clone())
    msgs = []
    if cur["sum_tt"] != Rn:
        msgs.append(f"sum(tiles_touched)={cur['sum_tt']} != R={Rn}")
    cnt = torch.bincount(order.clamp(0, P - 1), minlength=P)
    if int(cnt.max()) != 1 or int(cnt.min()) != 1:
        msgs.append(f"depth order is not a permutation: {int((cnt == 0).sum())} missing, {int((cnt > 1).sum())} duplicated")
    if ref is None:
        ref = cur
    else:
        for k in ("R", "sum_tt"):
            if cur[k] != ref[k]:
                msgs.append(f"{k}: {cur[k]} != first {ref[k]}")
        for k in ("order", "pl", "color", "ranges"):
            if cur[k].shape != ref[k].shape or not torch.equal(cur[k], ref[k]):
                n = int((cur[k] != ref[k]).sum()) if cur[k].shape == ref[k].shape else -1
                msgs.append(f"{k} differs from the first iteration in {n} entries")
    if msgs:
        bad += 1
        print(f"iter {it}: " + "; ".join(msgs), flush=True)
print(f"stress: {iters} iterations, {bad} bad, R={ref['R']}, visible={nvis}")

# phase 2: the bench's pattern -- forward + backward back to back, no host work in between
gC, gD, gA = [g.to(dev) for g in O.synth_upstream_grads(W, H, seed=1)]
seen = {}
first_color = None
for it in range(iters * 3):
    color, radii, depth, alpha, state = R.rasterize_forward_raw(*args, s)
    R.rasterize_backward_raw(state, *args, s, gC, gD, gA)
    seen[state.num_rendered] = seen.get(state.num_rendered, 0) + 1
    if it % 16 == 0:
        if first_color is None:
            first_color = color.clone()
        elif not torch.equal(color, first_color):
            bad += 1
            print(f"phase 2 iter {it}: image differs in {int((color != first_color).sum())} values", flush=True)
torch.cuda.synchronize()
print(f"phase 2: {iters * 3} fwd+bwd steps, distinct num_rendered: {seen}")
if len(seen) != 1:
    bad += 1
sys.exit(1 if bad else 0)
