"""One full training step at BASELINE config 3 (1M Gaussians, 1920x1080, SH degree 3) on the GPU box, the way reference
train.py:135-208 runs it -- raw parameters -> activations + assembly -> rasterizer -> L1 + D-SSIM loss -> backward ->
densification statistics (train.py:192-193) -> Adam on the 12 parameter groups -- in two variants over the SAME
rasterizer and loss kernels:

  fused      scgaussian_b200.model.render (one assembly kernel each way for the small arrays; the operator reads the SH
             coefficients in place and writes dL/dfeatures_* itself) + model.add_densification_stats (one launch, no
             host sync) + optim.step_all (one Adam launch)
  fused_assembled_sh   the same with the SH coefficients going through the assembled [P,16,3] copy (SCGR_SPLIT_SH=0)
  torch      the reference's chain of torch activations / cats, its boolean-mask statistics statements and its two
             torch.optim.Adam optimizers

CUDA events over back-to-back steps after warm-up.  `python tools/train_step_time.py` prints one JSON object (also to
gpurun_out/train_step.json); bench.py calls `measure()` for its `train_step` entry.
The hybrid model is a stand-in (70 % of the synthetic Gaussians parameterised as rays through the origin)."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

LR = {"_zval": 1.6e-4, "_features_dc": 2e-3, "_features_rest": 1e-4, "_opacity": 5.5e-2, "_scaling": 5.5e-3,
      "_rotation": 1.5e-3}


class _Obj:
    pass


def _make_pc(sc, n_ray, dev):
    P = int(sc["means3D"].shape[0])
    pc = _Obj()
    pc.active_sh_degree = pc.max_sh_degree = 3
    m = sc["means3D"]
    z = m[:n_ray].norm(dim=1, keepdim=True)
    par = lambda t: torch.nn.Parameter(t.clone().to(dev).contiguous())  # noqa: E731
    pc._rayo, pc._rayd, pc._zval = torch.zeros(n_ray, 3, device=dev), (m[:n_ray] / z).to(dev), par(z)
    pc.bg_xyz = par(m[n_ray:])
    lsc, rot, op, sh = sc["scales"].log(), sc["rotations"] * 1.3, torch.logit(sc["opacities"]), sc["shs"]
    pc._scaling, pc.bg_scaling = par(lsc[:n_ray]), par(lsc[n_ray:])
    pc._rotation, pc.bg_rotation = par(rot[:n_ray]), par(rot[n_ray:])
    pc._opacity, pc.bg_opacity = par(op[:n_ray]), par(op[n_ray:])
    pc._features_dc, pc.bg_features_dc = par(sh[:n_ray, :1]), par(sh[n_ray:, :1])
    pc._features_rest, pc.bg_features_rest = par(sh[:n_ray, 1:]), par(sh[n_ray:, 1:])
    pc.xyz_gradient_accum, pc.denom = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)
    pc.max_radii2D = torch.zeros(P, device=dev)
    return pc


def _groups(pc, prefix):
    names = {"_zval": "bg_xyz"} if prefix == "bg" else {}
    out = []
    for k, lr in LR.items():
        attr = k if prefix == "" else names.get(k, "bg" + k)
        out.append({"params": [getattr(pc, attr)], "lr": lr, "name": attr})
    return out


def measure(sc, cam_d, dev, W, H, steps=20, warm=5, variants=("fused", "fused_assembled_sh", "torch")):
    """sc / cam_d: scgaussian_b200.synthetic.synth_scene / make_camera outputs (CPU tensors).  Returns the dict the
    command-line tool prints."""
    import ctypes as C
    from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer, _lib, model, optim
    from scgaussian_b200.losses import photometric_loss
    F = torch.nn.functional
    P = int(sc["means3D"].shape[0])
    n_ray = int(P * 0.7)
    cam = _Obj()
    cam.image_height, cam.image_width = H, W
    cam.FoVx, cam.FoVy = 2 * math.atan(cam_d["tanfovx"]), 2 * math.atan(cam_d["tanfovy"])
    cam.world_view_transform, cam.full_proj_transform = cam_d["viewmatrix"].to(dev), cam_d["projmatrix"].to(dev)
    cam.camera_center = cam_d["campos"].to(dev)
    pipe = _Obj()
    pipe.debug = pipe.compute_cov3D_python = pipe.convert_SHs_python = False
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, device=dev)

    def render_torch(pc):
        xyz = torch.cat([pc._rayo + pc._rayd * pc._zval, pc.bg_xyz])
        scal = torch.cat([torch.exp(pc._scaling), torch.exp(pc.bg_scaling)])
        rot = torch.cat([F.normalize(pc._rotation), F.normalize(pc.bg_rotation)])
        opa = torch.cat([torch.sigmoid(pc._opacity), torch.sigmoid(pc.bg_opacity)])
        shs = torch.cat((torch.cat([pc._features_dc, pc.bg_features_dc]),
                         torch.cat([pc._features_rest, pc.bg_features_rest])), dim=1)
        ssp = torch.zeros_like(xyz, requires_grad=True) + 0
        ssp.retain_grad()
        rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam_d["tanfovx"],
                                           tanfovy=cam_d["tanfovy"], bg=bg, scale_modifier=1.0,
                                           viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                                           sh_degree=3, campos=cam.camera_center, prefiltered=False, debug=False)
        color, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(
            means3D=xyz, means2D=ssp, shs=shs, colors_precomp=None, opacities=opa, scales=scal, rotations=rot,
            cov3D_precomp=None)
        return {"render": color, "radii": radii, "viewspace_points": ssp, "visibility_filter": radii > 0}

    def run(variant):
        # "fused": the operator reads features_dc / features_rest in place (split SH layout, the default);
        # "fused_assembled_sh": the same passes through the assembled [P,16,3] copy (SCGR_SPLIT_SH=0), for comparison
        os.environ["SCGR_SPLIT_SH"] = "0" if variant == "fused_assembled_sh" else "1"
        fused = variant.startswith("fused")
        pc = _make_pc(sc, n_ray, dev)
        if fused:
            oa, ob = optim.Adam(_groups(pc, ""), lr=0.0, eps=1e-15), optim.Adam(_groups(pc, "bg"), lr=0.0, eps=1e-15)
        else:
            oa = torch.optim.Adam(_groups(pc, ""), lr=0.0, eps=1e-15)
            ob = torch.optim.Adam(_groups(pc, "bg"), lr=0.0, eps=1e-15)
        losses = []

        def step():
            out = model.render(cam, pc, pipe, bg) if fused else render_torch(pc)
            loss = photometric_loss(out["render"], gt, 0.2)
            loss.backward()
            # reference train.py:190-193 (every iteration while iteration < densify_until_iter = the whole default run)
            if fused:
                model.add_densification_stats(pc, out["viewspace_points"], None, out["radii"])
            else:
                vis, radii = out["visibility_filter"], out["radii"]
                pc.max_radii2D[vis] = torch.max(pc.max_radii2D[vis], radii[vis])
                pc.xyz_gradient_accum[vis] += torch.norm(out["viewspace_points"].grad[vis, :2], dim=-1, keepdim=True)
                pc.denom[vis] += 1
            if fused:
                optim.step_all(oa, ob)
            else:
                oa.step()
                ob.step()
            oa.zero_grad(set_to_none=True)
            ob.zero_grad(set_to_none=True)
            losses.append(loss.detach())

        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "loss_first": float(losses[0]), "loss_last": float(losses[-1])}
        if fused:
            # per-kernel times of 3 more steps (CUDA events bracketing every launch inside the library)
            lib = _lib.load()
            lib.scgr_profile_enable(1)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            names, msarr = (C.c_char_p * 512)(), (C.c_float * 512)()
            n = lib.scgr_profile_fetch(names, msarr, 512)
            lib.scgr_profile_enable(0)
            per = {}
            for i in range(max(n, 0)):
                per.setdefault(names[i].decode(), []).append(float(msarr[i]) * 1e3)
            res["kernel_us_per_step"] = {k: sum(v) / 3 for k, v in per.items()}
        return res

    res = {"P": P, "width": W, "height": H, "n_ray": n_ray, "n_bg": P - n_ray, "steps": steps, "warmup": warm}
    prev = os.environ.get("SCGR_SPLIT_SH")
    try:
        for v in variants:
            res[v] = run(v)
    finally:
        os.environ.pop("SCGR_SPLIT_SH", None)
        if prev is not None:
            os.environ["SCGR_SPLIT_SH"] = prev
    if "fused" in res and "torch" in res:
        res["speedup"] = res["torch"]["ms_per_step"] / res["fused"]["ms_per_step"]
    return res


def main():
    from scgaussian_b200 import synthetic as O  # SURVEY 8d seeded synthetic scene
    dev = torch.device("cuda", 0)
    P, W, H = int(os.environ.get("P", 1_000_000)), 1920, 1080
    variants = tuple(os.environ.get("VARIANTS", "fused,torch").split(","))
    res = measure(O.synth_scene(P, W, H, sh_degree=3, scale_median=0.01, seed=0), O.make_camera(W, H), dev, W, H,
                  steps=int(os.environ.get("STEPS", 20)), warm=int(os.environ.get("WARM", 5)), variants=variants)
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_step.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
