"""Forward+backward time at small scene sizes (GPU box): BASELINE config 2's shape (504x378, 30k-200k Gaussians)
through the raw stage calls and through the public operator + autograd."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scgaussian_b200 import synthetic as O
from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer
from scgaussian_b200 import rasterizer as R

dev = torch.device("cuda", 0)
for P, W, H, smed in [(30_000, 504, 378, 0.03), (200_000, 504, 378, 0.02), (1_000_000, 1920, 1080, 0.01)]:
    cam = O.make_camera(W, H)
    sc = O.synth_scene(P, W, H, sh_degree=3, scale_median=smed, seed=0)
    t = {k: v.to(dev).contiguous() for k, v in sc.items()}
    gC, gD, gA = [g.to(dev).contiguous() for g in O.synth_upstream_grads(W, H, seed=1)]
    s = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], torch.zeros(3, device=dev), 1.0,
                                      cam["viewmatrix"].to(dev), cam["projmatrix"].to(dev), 3, cam["campos"].to(dev), False, False)
    args_in = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None)

    def raw():
        color, radii, depth, alpha, state = R.rasterize_forward_raw(*args_in, s)
        R.rasterize_backward_raw(state, *args_in, s, gC, gD, gA)
        return state

    leaves = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)

    def op():
        color, radii, depth, alpha = GaussianRasterizer(s)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                          shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        ((color * gC).sum() + (depth * gD).sum() + (alpha * gA).sum()).backward()
        for v in leaves.values():
            v.grad = None
        m2d.grad = None

    for name, fn in (("raw", raw), ("operator+autograd", op)):
        for _ in range(10):
            st = fn()
        torch.cuda.synchronize()
        n = 100 if P < 500_000 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            st = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e6
        print(f"P={P} {W}x{H} {name}: {e0.elapsed_time(e1) / n * 1000:.1f} us/step (gpu events), {wall:.1f} us wall"
              + (f", R={st.num_rendered}" if st is not None else ""), flush=True)
