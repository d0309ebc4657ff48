"""Generates tests/golden/reference_anchors.npz by IMPORTING the reference's own Python
(/root/reference) -- the only reference-owned arithmetic on the rasterizer path (SURVEY.md
section 2a row 4, section 8c).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Vectors:
  eval_sh ................ reference utils/sh_utils.py:57-112, degrees 0..3, the layout
                           transform of reference gaussian_renderer/__init__.py:78-83
                           (colour = clamp_min(eval_sh + 0.5, 0))
  getProjectionMatrix .... reference utils/graphics_utils.py:51-71
  geom_transform_points .. reference utils/graphics_utils.py:22-29
  build_rotation / build_scaling_rotation / strip_symmetric
                           reference utils/general_utils.py:70-116 (run on CPU by mapping the
                           hard-coded device="cuda" literal to "cpu" for the duration of the call)
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from utils.sh_utils import eval_sh  # noqa: E402
from utils import graphics_utils as GU  # noqa: E402
from utils import general_utils as GenU  # noqa: E402


class _CpuZeros:
    """general_utils.py hard-codes device='cuda' in torch.zeros; redirect to CPU."""

    def __enter__(self):
        self._orig = torch.zeros

        def zeros(*a, **k):
            k.pop("device", None)
            return self._orig(*a, **k)
        torch.zeros = zeros

    def __exit__(self, *exc):
        torch.zeros = self._orig


def main():
    g = torch.Generator().manual_seed(1234)
    P = 64
    out = {}
    # --- SH -------------------------------------------------------------------------
    shs = torch.randn(P, 16, 3, generator=g, dtype=torch.float64) * 0.3      # [P,M,3] as the op gets it
    xyz = torch.randn(P, 3, generator=g, dtype=torch.float64) * 2.0
    campos = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64)
    out["sh_shs"], out["sh_xyz"], out["sh_campos"] = shs.numpy(), xyz.numpy(), campos.numpy()
    for deg in range(4):
        shs_view = shs.transpose(1, 2).view(-1, 3, 16)                       # reference __init__.py:79
        dir_pp = xyz - campos.repeat(P, 1)
        dirn = dir_pp / dir_pp.norm(dim=1, keepdim=True)
        sh2rgb = eval_sh(deg, shs_view, dirn)
        out[f"sh_rgb_deg{deg}"] = torch.clamp_min(sh2rgb + 0.5, 0.0).numpy()  # reference __init__.py:83
    # --- projection -----------------------------------------------------------------
    fovx, fovy = 1.0471975511965976, 0.6435011087932844
    Pm = GU.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy)
    out["proj_fov"] = np.array([fovx, fovy])
    out["proj_matrix"] = Pm.numpy()
    Rm = np.array([[0.9, -0.1, 0.42], [0.15, 0.98, -0.08], [-0.4, 0.13, 0.9]])
    Rm, _ = np.linalg.qr(Rm)
    tv = np.array([0.1, -0.3, 0.7])
    w2v = torch.tensor(GU.getWorld2View2(Rm, tv)).transpose(0, 1)           # reference cameras.py:60
    full = (w2v.unsqueeze(0).bmm(Pm.transpose(0, 1).unsqueeze(0))).squeeze(0)  # cameras.py:61-62
    pts = torch.randn(P, 3, generator=g) * 1.5 + torch.tensor([0.0, 0.0, 4.0])
    out["xf_R"], out["xf_t"] = Rm, tv
    out["xf_world_view"] = w2v.numpy()
    out["xf_full_proj"] = full.numpy()
    out["xf_campos"] = w2v.inverse()[3, :3].numpy()                           # cameras.py:63
    out["xf_points"] = pts.numpy()
    out["xf_ndc"] = GU.geom_transform_points(pts, full).numpy()
    # --- covariance -----------------------------------------------------------------
    scal = torch.exp(torch.randn(P, 3, generator=g) * 0.5 - 3.0)
    rot = torch.randn(P, 4, generator=g)          # NOT normalised: build_rotation normalises itself
    with _CpuZeros():
        R = GenU.build_rotation(rot)
        L = GenU.build_scaling_rotation(1.7 * scal, rot)
        cov6 = GenU.strip_symmetric(L @ L.transpose(1, 2))                    # gaussian_model.py:37-41
    out["cov_scales"], out["cov_rots"] = scal.numpy(), rot.numpy()
    out["cov_modifier"] = np.array(1.7)
    out["cov_R"], out["cov_sixvec"] = R.numpy(), cov6.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_anchors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
