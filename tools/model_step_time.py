"""Times the two elementwise passes around the rasterizer on the GPU box at BASELINE
config 3's model size (1M Gaussians, SH degree 3; 70 % ray-based + 30 % free as a stand-in for the hybrid model):

  * fused assembly forward / backward (scgr_assemble_*) against the reference's chain of torch activations + cats and
    its autograd backward (reference scene/gaussian_model.py:105-152);
  * fused Adam over the reference's 12 parameter groups (scgr_adam_step, one launch) against the reference's two
    torch.optim.Adam optimizers (reference scene/gaussian_model.py:491-512, train.py:204-208).

CUDA events around back-to-back iterations after warm-up (each pass streams > 400 MB: larger than the 126 MB L2);
algorithmic bytes from scgaussian_b200/csrc/model.cu's header.  Prints one JSON object (also to gpurun_out/).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from scgaussian_b200 import model, optim  # noqa: E402

dev = torch.device("cuda", 0)
P = int(os.environ.get("P", 1_000_000))
n_ray, K = int(P * 0.7), 16
n_bg = P - n_ray
g = torch.Generator(device=dev).manual_seed(0)
F = torch.nn.functional


def par(*s, shift=0.0):
    return torch.nn.Parameter(torch.randn(*s, generator=g, device=dev) + shift)


raw = dict(rayo=torch.randn(n_ray, 3, generator=g, device=dev), rayd=torch.randn(n_ray, 3, generator=g, device=dev),
           zval=par(n_ray, 1, shift=3), scaling=par(n_ray, 3, shift=-4), rotation=par(n_ray, 4), opacity=par(n_ray, 1),
           features_dc=par(n_ray, 1, 3), features_rest=par(n_ray, K - 1, 3),
           bg_xyz=par(n_bg, 3), bg_scaling=par(n_bg, 3, shift=-4), bg_rotation=par(n_bg, 4), bg_opacity=par(n_bg, 1),
           bg_features_dc=par(n_bg, 1, 3), bg_features_rest=par(n_bg, K - 1, 3))
trained = [k for k in raw if k not in ("rayo", "rayd")]
ups = None


def reference_chain():
    r = raw
    xyz = torch.cat([r["rayo"] + r["rayd"] * r["zval"], r["bg_xyz"]])
    scal = torch.cat([torch.exp(r["scaling"]), torch.exp(r["bg_scaling"])])
    rot = torch.cat([F.normalize(r["rotation"]), F.normalize(r["bg_rotation"])])
    opa = torch.cat([torch.sigmoid(r["opacity"]), torch.sigmoid(r["bg_opacity"])])
    shs = torch.cat((torch.cat([r["features_dc"], r["bg_features_dc"]]),
                     torch.cat([r["features_rest"], r["bg_features_rest"]])), dim=1)
    return xyz, scal, rot, opa, shs


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3      # us


def fwd_bwd(assemble_fn):
    global ups
    outs = assemble_fn()
    if ups is None:
        ups = [torch.randn_like(o) for o in outs]
    torch.autograd.backward(outs, ups)
    for k in trained:
        raw[k].grad = None


res = {"P": P, "n_ray": n_ray, "n_bg": n_bg, "sh_coeffs": K}
with torch.no_grad():
    res["assemble_forward_us"] = timed(lambda: model.assemble(**raw))
    res["torch_forward_us"] = timed(reference_chain)
res["assemble_fwd_bwd_us"] = timed(lambda: fwd_bwd(lambda: model.assemble(**raw)))
res["torch_fwd_bwd_us"] = timed(lambda: fwd_bwd(reference_chain))
res["assemble_backward_us"] = res["assemble_fwd_bwd_us"] - res["assemble_forward_us"]
fwd_bytes = n_ray * (252 + 236) + n_bg * (240 + 236)
bwd_bytes = P * (236 + 44 + 228)
res["assemble_forward_gbs"] = fwd_bytes / res["assemble_forward_us"] / 1e3
res["assemble_backward_gbs_upper"] = bwd_bytes / max(res["assemble_backward_us"], 1e-3) / 1e3

# the optimizer step: the reference's 12 groups
lr = {"zval": 4e-4, "xyz": 4e-4, "features_dc": 2e-3, "features_rest": 1e-4, "opacity": 5.5e-2, "scaling": 5.5e-3,
      "rotation": 1.5e-3}
for k in trained:
    raw[k].grad = torch.randn_like(raw[k]) * 1e-3
groups = [{"params": [raw[k]], "lr": lr[k.replace("bg_", "")], "name": k} for k in trained]
main = [gr for gr in groups if not gr["name"].startswith("bg_")]
bg = [gr for gr in groups if gr["name"].startswith("bg_")]
ours_a, ours_b = optim.Adam(main, lr=0.0, eps=1e-15), optim.Adam(bg, lr=0.0, eps=1e-15)
res["adam_fused_us"] = timed(lambda: optim.step_all(ours_a, ours_b))
ref_a = torch.optim.Adam([dict(gr) for gr in main], lr=0.0, eps=1e-15)
ref_b = torch.optim.Adam([dict(gr) for gr in bg], lr=0.0, eps=1e-15)
res["adam_torch_us"] = timed(lambda: (ref_a.step(), ref_b.step()))
n_el = sum(raw[k].numel() for k in trained)
res["adam_elements"] = n_el
res["adam_fused_gbs"] = n_el * 28 / res["adam_fused_us"] / 1e3
# per-kernel times of the fused passes (CUDA events bracketing each launch inside the library)
import ctypes as C  # noqa: E402
from scgaussian_b200 import _lib  # noqa: E402
lib = _lib.load()
lib.scgr_profile_enable(1)
for _ in range(10):
    fwd_bwd(lambda: model.assemble(**raw))
    for k in trained:
        raw[k].grad = torch.zeros_like(raw[k])
    optim.step_all(ours_a, ours_b)
torch.cuda.synchronize()
names, msarr = (C.c_char_p * 256)(), (C.c_float * 256)()
n = lib.scgr_profile_fetch(names, msarr, 256)
lib.scgr_profile_enable(0)
per = {}
for i in range(max(n, 0)):
    per.setdefault(names[i].decode(), []).append(float(msarr[i]) * 1e3)
res["kernel_us_median"] = {k: sorted(v)[len(v) // 2] for k, v in per.items()}
res["assemble_backward_kernel_gbs"] = bwd_bytes / res["kernel_us_median"]["assemble_backward"] / 1e3
res["assemble_forward_kernel_gbs"] = fwd_bytes / res["kernel_us_median"]["assemble_forward"] / 1e3
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
res["hbm_peak_gbs"] = peaks["hbm_gbs"]
res["assemble_forward_frac"] = res["assemble_forward_gbs"] / peaks["hbm_gbs"]
res["adam_fused_frac"] = res["adam_fused_gbs"] / peaks["hbm_gbs"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
res["sh_copy"] = "flat" if os.environ.get("SCGR_ASSEMBLE_STAGED") == "0" else "staged"
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "model_time.json")), "w"), indent=1)
print(json.dumps(res))
