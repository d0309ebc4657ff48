"""Launches each of the model-pass kernels (assemble_forward, assemble_backward, adam)
a few times at 1M Gaussians / SH degree 3 for an ncu capture:

  ncu --set full --clock-control none --import-source on -k regex:"assemble_|adam_" -s 6 -c 3 \\
      -o gpurun_out/prof_model -f python tools/model_ncu.py
  python tools/summarize_ncu.py r01_model gpurun_out/prof_model.ncu-rep
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from scgaussian_b200 import model, optim  # noqa: E402

dev = torch.device("cuda", 0)
P, K = 1_000_000, 16
n_ray = int(P * 0.7)
n_bg = P - n_ray
g = torch.Generator(device=dev).manual_seed(0)
par = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g, device=dev))  # noqa: E731
raw = dict(rayo=torch.randn(n_ray, 3, generator=g, device=dev), rayd=torch.randn(n_ray, 3, generator=g, device=dev),
           zval=par(n_ray, 1), scaling=par(n_ray, 3), rotation=par(n_ray, 4), opacity=par(n_ray, 1),
           features_dc=par(n_ray, 1, 3), features_rest=par(n_ray, K - 1, 3),
           bg_xyz=par(n_bg, 3), bg_scaling=par(n_bg, 3), bg_rotation=par(n_bg, 4), bg_opacity=par(n_bg, 1),
           bg_features_dc=par(n_bg, 1, 3), bg_features_rest=par(n_bg, K - 1, 3))
trained = [k for k in raw if k not in ("rayo", "rayd")]
opt = optim.Adam([{"params": [raw[k]], "lr": 1e-3, "name": k} for k in trained], lr=0.0, eps=1e-15)
ups = None
for _ in range(3):          # launches 1-3, 4-6 warm up; ncu captures the third round (-s 6 -c 3)
    outs = model.assemble(**raw)
    ups = ups or [torch.randn_like(o) for o in outs]
    torch.autograd.backward(outs, ups)
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
