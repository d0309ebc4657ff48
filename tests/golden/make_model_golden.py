"""Generates tests/golden/model_golden.npz by IMPORTING the reference's own model code
(/root/reference/scene/gaussian_model.py) and running it on a small seeded hybrid model on the CPU:

  * the activation / assembly properties get_xyz :123-129, get_scaling :105-113, get_rotation :115-122,
    get_opacity :142-150, get_features :131-140 and their autograd gradients for a random linear functional
    (SURVEY.md section 8f row f2);
  * the two optimizers exactly as `training_setup` builds them (:486-512: torch.optim.Adam(l, lr=0.0, eps=1e-15),
    learning rates from arguments/__init__.py OptimizationParams) stepped 3 times with `update_learning_rate`
    (:514-527) and seeded gradients (row f3);
  * the per-iteration statistics update of train.py:192-193 (the statement at :192 is read from the file and executed
    verbatim; :193 calls GaussianModel.add_densification_stats :932-934) (row a17).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_model_golden.py

The reference's unrelated, uninstalled imports (plyfile, simple_knn, pytorch3d, skimage, imageio, matplotlib, dkm, lpips) are stubbed; `training_setup`
allocates two statistics tensors with device="cuda" (:488-489), which is redirected to the CPU for the run.
"""
import argparse
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
_STUB_ROOTS = ("plyfile", "simple_knn", "pytorch3d", "skimage", "imageio", "matplotlib", "dkm", "lpips")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def build_model(GaussianModel, g, n_ray, n_bg, sh_degree=3):
    """A hybrid model with the attribute names / shapes the reference creates (:398-470)."""
    K = (sh_degree + 1) ** 2
    pc = GaussianModel(sh_degree)
    P = torch.nn.Parameter

    def rnd(*s, scale=1.0):
        return torch.randn(*s, generator=g) * scale

    pc._rayo = rnd(n_ray, 3)
    pc._rayd = torch.nn.functional.normalize(rnd(n_ray, 3))
    pc._zval = P(torch.rand(n_ray, 1, generator=g) * 4 + 1)
    pc._features_dc = P(rnd(n_ray, 1, 3, scale=0.5))
    pc._features_rest = P(rnd(n_ray, K - 1, 3, scale=0.1))
    pc._scaling = P(rnd(n_ray, 3) - 3.0)
    pc._rotation = P(rnd(n_ray, 4))
    pc._opacity = P(rnd(n_ray, 1, scale=2.0))
    pc.bg_xyz = P(rnd(n_bg, 3, scale=3.0))
    pc.bg_features_dc = P(rnd(n_bg, 1, 3, scale=0.5))
    pc.bg_features_rest = P(rnd(n_bg, K - 1, 3, scale=0.1))
    pc.bg_scaling = P(rnd(n_bg, 3) - 3.0)
    pc.bg_rotation = P(rnd(n_bg, 4))
    pc.bg_opacity = P(rnd(n_bg, 1, scale=2.0))
    return pc


RAW = ("_rayo", "_rayd", "_zval", "_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest",
       "bg_xyz", "bg_scaling", "bg_rotation", "bg_opacity", "bg_features_dc", "bg_features_rest")
TRAINED = tuple(n for n in RAW if n not in ("_rayo", "_rayd"))
ACT = ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features")


def main():
    finder = _StubFinder()
    sys.meta_path.append(finder)
    sys.path.insert(0, REF)
    from scene.gaussian_model import GaussianModel
    from arguments import OptimizationParams

    out = {}
    g = torch.Generator().manual_seed(11)
    # ---- f2: activations + assembly, two shapes (hybrid; ray-based only as before the bg set exists)
    for tag, (n_ray, n_bg) in {"hyb": (37, 21), "ray": (19, 0)}.items():
        pc = build_model(GaussianModel, g, n_ray, n_bg)
        acts = [getattr(pc, a) for a in ACT]
        ws = [torch.randn(a.shape, generator=g) for a in acts]
        loss = sum((a * w).sum() for a, w in zip(acts, ws))
        params = [getattr(pc, n) for n in TRAINED if getattr(pc, n).shape[0] > 0]
        names = [n for n in TRAINED if getattr(pc, n).shape[0] > 0]
        grads = torch.autograd.grad(loss, params)
        for n in RAW:
            out[f"{tag}_raw{n}"] = getattr(pc, n).detach().numpy()
        for a, t, w in zip(ACT, acts, ws):
            out[f"{tag}_{a}"] = t.detach().numpy()
            out[f"{tag}_w_{a}"] = w.numpy()
        for n, gr in zip(names, grads):
            out[f"{tag}_grad{n}"] = gr.numpy()

    # ---- f3: the reference's two Adam optimizers, 3 steps
    pc = build_model(GaussianModel, g, 23, 9)
    opt_args = OptimizationParams(argparse.ArgumentParser())
    pc.spatial_lr_scale = 2.5
    real_zeros = torch.zeros
    torch.zeros = lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    try:
        pc.training_setup(opt_args)
    finally:
        torch.zeros = real_zeros
    for n in TRAINED:
        out[f"adam_p0{n}"] = getattr(pc, n).detach().numpy().copy()
    lrs = {}
    for it in (1, 2, 3):
        pc.update_learning_rate(it)
        for n in TRAINED:
            p = getattr(pc, n)
            p.grad = torch.randn(p.shape, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
            out[f"adam_g{it}{n}"] = p.grad.numpy().copy()
        pc.optimizer.step()
        pc.optimizer_bg.step()
        for opt in (pc.optimizer, pc.optimizer_bg):
            for grp in opt.param_groups:
                lrs[(it, grp["name"])] = grp["lr"]
        for n in TRAINED:
            out[f"adam_p{it}{n}"] = getattr(pc, n).detach().numpy().copy()
    group_of = {"_zval": "zval", "_features_dc": "f_dc", "_features_rest": "f_rest", "_opacity": "opacity",
                "_scaling": "scaling", "_rotation": "rotation", "bg_xyz": "bg_xyz", "bg_features_dc": "bg_f_dc",
                "bg_features_rest": "bg_f_rest", "bg_opacity": "bg_opacity", "bg_scaling": "bg_scaling",
                "bg_rotation": "bg_rotation"}
    for n in TRAINED:
        out[f"adam_lr{n}"] = np.array([lrs[(it, group_of[n])] for it in (1, 2, 3)], dtype=np.float64)
        opt = pc.optimizer if not n.startswith("bg_") else pc.optimizer_bg
        st = opt.state[getattr(pc, n)]
        out[f"adam_m3{n}"] = st["exp_avg"].numpy().copy()
        out[f"adam_v3{n}"] = st["exp_avg_sq"].numpy().copy()
    out["adam_eps"] = np.float64(pc.optimizer.defaults["eps"])
    out["adam_betas"] = np.array(pc.optimizer.defaults["betas"], dtype=np.float64)

    # ---- a17: the per-iteration densification statistics (train.py:192-193 -> gaussian_model.py:932-934)
    n = 57
    pc = GaussianModel(3)
    pc.xyz_gradient_accum = torch.rand(n, 1, generator=g)
    pc.denom = torch.randint(0, 5, (n, 1), generator=g).float()
    pc.max_radii2D = torch.randint(0, 30, (n,), generator=g).float()
    radii = torch.randint(-1, 40, (n,), generator=g).to(torch.int32).clamp_min(0)
    radii[::5] = 0
    vsp = torch.zeros(n, 3, requires_grad=True)
    vsp.grad = torch.randn(n, 3, generator=g) * 1e-3
    out["stats_accum0"], out["stats_denom0"] = pc.xyz_gradient_accum.numpy().copy(), pc.denom.numpy().copy()
    out["stats_maxr0"], out["stats_radii"], out["stats_grad"] = pc.max_radii2D.numpy().copy(), radii.numpy(), vsp.grad.numpy()
    lines = open(os.path.join(REF, "train.py")).read().splitlines()
    stmt = [ln.strip() for ln in lines if ln.strip().startswith("gaussians.max_radii2D[visibility_filter] =")]
    assert len(stmt) == 1, stmt                      # train.py:192, executed verbatim
    scope = {"gaussians": pc, "visibility_filter": radii > 0, "radii": radii, "torch": torch}
    exec(stmt[0], scope)
    pc.add_densification_stats(vsp, radii > 0)       # train.py:193 -> gaussian_model.py:932-934
    out["stats_accum1"], out["stats_denom1"] = pc.xyz_gradient_accum.numpy().copy(), pc.denom.numpy().copy()
    out["stats_maxr1"] = pc.max_radii2D.numpy().copy()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays,", os.path.getsize(path), "bytes")
    print({k: v for k, v in lrs.items() if k[0] == 3})


if __name__ == "__main__":
    main()
