"""Generates tests/golden/prior_golden.npz by IMPORTING / EXECUTING the reference's own code for BASELINE config 5's
loss path, on the CPU:

  * `GaussianModel.get_matchloss_from_renderdepth` (/root/reference/scene/gaussian_model.py:241-282), unmodified, on a
    seeded 3-view `view_gs` (intrinsics, world-to-camera matrices, matched pixels, rays built with the formulas of
    `create_from_mono` :300-336) and a rendered-depth image, with its autograd gradient with respect to the depth;
  * the DTU background statements of /root/reference/train.py:151-158 and the alpha term :167-168 -- the source lines
    are read from the file and executed verbatim on a seeded ground-truth image and alpha map.

Run in the build container only (the GPU box has no /root/reference):   python tests/golden/make_prior_golden.py
"""
import importlib.abc
import importlib.machinery
import math
import os
import sys
import textwrap
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


def make_views(g, n_views=3, W=96, H=72, M=400):
    """A seeded multi-view setup in the layout of GaussianModel.view_gs (reference :284-366): a bumpy surface in front of
    slightly rotated cameras; matches = projections of common surface points, perturbed by a pixel or two."""
    fx = fy = 0.9 * W
    intr = torch.tensor([[fx, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1]]).float()
    views, names = {}, [f"view{k}" for k in range(n_views)]
    w2cs = []
    for k in range(n_views):
        a = math.radians(6.0 * (k - 1))
        Rm = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]]).float()
        w2c = torch.eye(4)
        w2c[:3, :3] = Rm
        w2c[:3, 3] = torch.tensor([0.15 * (k - 1), 0.02 * k, 0.1 * k])
        w2cs.append(w2c)
    for k, name in enumerate(names):
        views[name] = {"intr": intr.clone(), "w2c": w2cs[k], "width": W, "height": H, "match_infos": {}}
    for i in range(n_views):
        for j in range(n_views):
            if i == j:
                continue
            gg = torch.Generator().manual_seed(100 * min(i, j) + max(i, j))      # the same 3D points for (i, j) and (j, i)
            pts = torch.stack([(torch.rand(M, generator=gg) - 0.5) * 3.0, (torch.rand(M, generator=gg) - 0.5) * 2.2,
                               4.0 + torch.rand(M, generator=gg) * 2.0], -1)
            cam = (w2cs[i] @ torch.cat([pts, torch.ones(M, 1)], 1).T)[:3]
            px = (intr @ cam)
            uv = (px[:2] / px[2:]).T + torch.randn(M, 2, generator=g) * 0.7                 # (M, 2) pixels
            # rays exactly as create_from_mono builds them (:330-336)
            homo = torch.cat([uv, torch.ones(M, 1)], 1)
            p = torch.matmul(intr.inverse()[None, :3, :3], homo[:, :, None]).squeeze()
            rays_d = p / (torch.linalg.norm(p, ord=2, dim=-1, keepdim=True) + 1e-8)
            rays_d = torch.matmul(w2cs[i].inverse()[None, :3, :3], rays_d[:, :, None]).squeeze()
            rays_o = w2cs[i].inverse()[None, :3, 3].expand(rays_d.shape).contiguous()
            cam_rays_d = torch.matmul(w2cs[i][None, :3, :3], rays_d[:, :, None]).squeeze()
            mask = (torch.rand(M, generator=g) > 0.15).float() * torch.rand(M, generator=g)     # warp_mask: bilinear samples in [0, 1]
            views[names[i]]["match_infos"][names[j]] = {"uv": uv, "rays_o": rays_o, "rays_d": rays_d, "cam_rays_d": cam_rays_d,
                                                        "blender_mask": mask}
    return views, names


def main():
    finder = _StubFinder()
    sys.meta_path.append(finder)
    sys.path.insert(0, REF)
    from scene.gaussian_model import GaussianModel
    out = {}
    g = torch.Generator().manual_seed(21)
    views, names = make_views(g)
    pc = GaussianModel(3)
    pc.view_gs = views
    W, H = views[names[0]]["width"], views[names[0]]["height"]
    for vi, name in enumerate(names):
        cam0 = types.SimpleNamespace(intr=views[name]["intr"], w2c=views[name]["w2c"], image_name=name)
        yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        depth = (5.0 + 0.6 * torch.sin(xx / 9.0 + vi) * torch.cos(yy / 7.0) + 0.05 * torch.randn(H, W, generator=g)).unsqueeze(0)
        depth.requires_grad_(True)
        loss = pc.get_matchloss_from_renderdepth(cam0, depth, None)
        grad, = torch.autograd.grad(loss, depth)
        out[f"match_depth_{vi}"], out[f"match_loss_{vi}"], out[f"match_grad_{vi}"] = depth.detach().numpy(), loss.item(), grad.numpy()
    out["match_size"] = np.array([W, H])
    out["match_names"] = np.array(names)
    for name in names:
        out[f"view_{name}_intr"], out[f"view_{name}_w2c"] = views[name]["intr"].numpy(), views[name]["w2c"].numpy()
        for other, md in views[name]["match_infos"].items():
            for k, v in md.items():
                out[f"view_{name}_{other}_{k}"] = v.numpy()

    # ---- reference train.py:151-158 + :167-168, executed verbatim ----
    src = open(os.path.join(REF, "train.py")).read().splitlines()
    block = "\n".join(src[150:158])          # lines 151..158: `if 'scan110' not in ...` ... `gt_image[bg_mask.repeat(3,1,1)] = 0.`
    assert "bg_mask_clone = bg_mask.clone()" in block and "for i in range(1, 50):" in block, block
    alpha_line = src[167].strip()            # line 168: loss += render_pkg["rendered_alpha"][bg_mask].mean()
    assert alpha_line == 'loss += render_pkg["rendered_alpha"][bg_mask].mean()', alpha_line
    for tag, (h, w) in {"a": (83, 61), "b": (150, 40)}.items():
        gt = torch.rand(3, h, w, generator=g) * 0.5
        # dark regions of assorted heights: runs shorter and longer than the 50-row window, one touching the top edge
        gt[:, : h // 2, : w // 3] *= 0.1
        gt[:, h // 3: h // 3 + 30, w // 2:] *= 0.05
        gt[:, 5:, w // 3: w // 2] *= 0.15
        gt_image = gt.clone()
        alpha = torch.rand(1, h, w, generator=g, requires_grad=True)
        env = {"gt_image": gt_image, "args": types.SimpleNamespace(source_path="/data/dtu/scan30"), "torch": torch}
        exec(textwrap.dedent(block), env)
        bg_mask = env["bg_mask"]
        env2 = {"loss": torch.zeros(()), "render_pkg": {"rendered_alpha": alpha}, "bg_mask": bg_mask}
        exec(alpha_line, env2)
        ga, = torch.autograd.grad(env2["loss"], alpha)
        out[f"bg_{tag}_gt"], out[f"bg_{tag}_gt_masked"], out[f"bg_{tag}_mask"] = gt.numpy(), env["gt_image"].numpy(), bg_mask.numpy()
        out[f"bg_{tag}_alpha"], out[f"bg_{tag}_alpha_mean"], out[f"bg_{tag}_alpha_grad"] = alpha.detach().numpy(), env2["loss"].item(), ga.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prior_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: out[k] for k in out if k.startswith("match_loss") or k.endswith("alpha_mean")},
          {k: int(out[k].sum()) for k in out if k.endswith("_mask")})


if __name__ == "__main__":
    main()
