"""The two elementwise passes around the rasterizer in a training step:

* activations + hybrid assembly = reference scene/gaussian_model.py:105-152 (get_xyz / get_scaling / get_rotation /
  get_opacity / get_features), fused forward + backward (scgaussian_b200/model.py -> scgr_assemble_*);
* the optimizer step = torch.optim.Adam(l, lr=0.0, eps=1e-15) as reference scene/gaussian_model.py:486-512 builds it
  and train.py:204-208 steps it (scgaussian_b200/optim.py -> scgr_adam_step).

CPU: the oracle (oracle/model_oracle.py) is PINNED to the reference's own code through tests/golden/
model_golden.npz (tests/golden/make_model_golden.py imports /root/reference/scene/gaussian_model.py).  GPU: the CUDA
kernels (scgaussian_b200/csrc/model.cu, through the C ABI) against the golden vectors, the oracle and torch's own
optimizer.  Tolerances: activations are single fp32 operations (exp, sigmoid, divide) -> 2e-6 relative;
gradients 1e-5; Adam 1e-6 of the parameter scale per step (a few ulp: torch's CUDA kernels contract a*b+c
differently from its CPU ones as well)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as MO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_golden.npz")
ACT = ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features")
RAY_KEYS = ("rayo", "rayd", "zval", "scaling", "rotation", "opacity", "features_dc", "features_rest")
BG_KEYS = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
TRAINED = ("_zval", "_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest",
           "bg_xyz", "bg_scaling", "bg_rotation", "bg_opacity", "bg_features_dc", "bg_features_rest")
RTOL_ACT = 2e-6
RTOL_GRAD = 1e-5


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _golden_sets(g, tag):
    ray = {k: g[f"{tag}_raw_{k}"] for k in RAY_KEYS}
    bg = {k: g[f"{tag}_rawbg_{k}"] for k in BG_KEYS}
    return (ray if ray["zval"].shape[0] else {}), (bg if bg["xyz"].shape[0] else {})


def _leaf_name(trained):      # "_zval" -> "ray_zval", "bg_xyz" -> "bg_xyz"
    return "ray" + trained if trained.startswith("_") else trained


# ------------------------------------------------------------------------------------------ CPU: oracle pinned
@pytest.mark.parametrize("tag", ["hyb", "ray"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_golden(tag, dtype):
    g = np.load(GOLD)
    ray, bg = _golden_sets(g, tag)
    acts, leaves = MO.assemble(ray, bg, dtype)
    for name, t in zip(ACT, acts):
        assert t.shape == g[f"{tag}_{name}"].shape
        assert _rel(t.detach().numpy(), g[f"{tag}_{name}"]) < RTOL_ACT, name
    loss = sum((a * torch.as_tensor(g[f"{tag}_w_{n}"]).to(dtype)).sum() for n, a in zip(ACT, acts))
    names = [n for n in TRAINED if f"{tag}_grad{n}" in g.files]
    assert names, "golden file holds no gradients"
    grads = torch.autograd.grad(loss, [leaves[_leaf_name(n)] for n in names])
    for n, gr in zip(names, grads):
        assert _rel(gr.numpy(), g[f"{tag}_grad{n}"]) < RTOL_GRAD, n


def test_adam_oracle_matches_reference_golden():
    g = np.load(GOLD)
    beta1, beta2 = g["adam_betas"]
    eps = float(g["adam_eps"])
    assert eps == 1e-15 and (beta1, beta2) == (0.9, 0.999)        # reference scene/gaussian_model.py:502
    for n in TRAINED:
        p = g[f"adam_p0{n}"]
        m, v = np.zeros_like(p), np.zeros_like(p)
        for it in (1, 2, 3):
            p, m, v = MO.adam_step(p, g[f"adam_g{it}{n}"], m, v, it, float(g[f"adam_lr{n}"][it - 1]), beta1, beta2, eps)
            assert np.abs(p - g[f"adam_p{it}{n}"]).max() <= 1e-6 * np.abs(p).max(), (n, it)
        assert _rel(m, g[f"adam_m3{n}"]) < 1e-6 and _rel(v, g[f"adam_v3{n}"]) < 1e-6, n
    # the reference's scheduler quirk is part of the pinned behaviour: update_learning_rate returns from inside the
    # first loop (:519), so bg_xyz keeps its initial rate while zval decays
    assert g["adam_lr_zval"][2] < g["adam_lr_zval"][0] and g["adam_lrbg_xyz"][2] == g["adam_lrbg_xyz"][0]


def test_stats_oracle_matches_reference_golden():
    g = np.load(GOLD)
    radii = g["stats_radii"]
    a, d, m = MO.densification_stats(g["stats_grad"], radii > 0, radii, g["stats_accum0"], g["stats_denom0"], g["stats_maxr0"])
    assert np.array_equal(d, g["stats_denom1"]) and np.array_equal(m, g["stats_maxr1"])
    assert _rel(a, g["stats_accum1"]) < 1e-6
    assert (radii == 0).any() and (g["stats_maxr1"] != g["stats_maxr0"]).any()


# ------------------------------------------------------------------------------------------ CPU: host logic
def test_no_cpu_path():
    from scgaussian_b200 import model, optim
    from scgaussian_b200._lib import ScgrError
    n = 4
    with pytest.raises(ScgrError, match="CUDA"):
        model.assemble(bg_xyz=torch.zeros(n, 3), bg_scaling=torch.zeros(n, 3), bg_rotation=torch.ones(n, 4),
                       bg_opacity=torch.zeros(n, 1), bg_features_dc=torch.zeros(n, 1, 3),
                       bg_features_rest=torch.zeros(n, 15, 3))
    with pytest.raises(ScgrError, match="no Gaussians"):
        model.assemble()
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.ones(5)
    opt = optim.Adam([{"params": [p], "lr": 0.1, "name": "x"}], lr=0.0, eps=1e-15)
    with pytest.raises(ScgrError, match="CUDA"):
        opt.step()
    with pytest.raises(ScgrError):
        optim.Adam([p], weight_decay=0.1)
    with pytest.raises(ScgrError):
        optim.Adam([p], amsgrad=True)


def test_fused_adam_keeps_torch_optimizer_surface():
    """What the reference does to its optimizers besides step(): per-group lr scheduling by name
    (scene/gaussian_model.py:514-527), state surgery by parameter (:758-843), state_dict round trips (:83, :103)."""
    from scgaussian_b200 import optim
    a, b = torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(7, 1))
    groups = [{"params": [a], "lr": 0.5, "name": "zval"}, {"params": [b], "lr": 0.25, "name": "opacity"}]
    ref = torch.optim.Adam([dict(g) for g in groups], lr=0.0, eps=1e-15)
    a.grad, b.grad = torch.randn_like(a), torch.randn_like(b)
    ref.step()                                              # torch fills the state on the CPU
    ours = optim.Adam([dict(g) for g in groups], lr=0.0, eps=1e-15)
    assert [g["name"] for g in ours.param_groups] == ["zval", "opacity"] and ours.defaults["eps"] == 1e-15
    ours.load_state_dict(ref.state_dict())                  # torch's checkpoint loads into the fused optimizer
    st = ours.state[a]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0
    assert torch.equal(st["exp_avg"], ref.state[a]["exp_avg"])
    back = torch.optim.Adam([dict(g) for g in groups], lr=0.0, eps=1e-15)
    back.load_state_dict(ours.state_dict())                 # ... and back
    assert torch.equal(back.state[b]["exp_avg_sq"], ref.state[b]["exp_avg_sq"])
    for grp in ours.param_groups:
        if grp["name"] == "zval":
            grp["lr"] = 0.125
    assert ours.param_groups[0]["lr"] == 0.125
    ours.zero_grad(set_to_none=True)
    assert a.grad is None
    ours.step()                                             # nothing has a gradient: nothing to launch, no error


# ------------------------------------------------------------------------------------------ GPU: parity
def _cuda_sets(ray, bg, requires_grad=True):
    def conv(d, trained):
        out = {}
        for k, v in d.items():
            t = torch.tensor(np.asarray(v), dtype=torch.float32, device="cuda")
            out[k] = t.requires_grad_(requires_grad and k in trained)
        return out
    return conv(ray, RAY_KEYS[2:]), conv(bg, BG_KEYS)


def _run_fused(ray, bg):
    from scgaussian_b200 import model
    kw = dict(ray)
    kw.update({"bg_" + k: v for k, v in bg.items()})
    return model.assemble(**kw)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["hyb", "ray"])
def test_assemble_matches_reference_golden(tag):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    g = np.load(GOLD)
    ray, bg = _golden_sets(g, tag)
    cray, cbg = _cuda_sets(ray, bg)
    acts = _run_fused(cray, cbg)
    for name, t in zip(ACT, acts):
        assert tuple(t.shape) == g[f"{tag}_{name}"].shape
        assert _rel(t.detach().cpu().numpy(), g[f"{tag}_{name}"]) < RTOL_ACT, name
    loss = sum((a * torch.tensor(g[f"{tag}_w_{n}"], device="cuda")).sum() for n, a in zip(ACT, acts))
    loss.backward()
    for n in TRAINED:
        if f"{tag}_grad{n}" not in g.files:
            continue
        t = cray[n[1:]] if n.startswith("_") else cbg[n[3:]]
        assert _rel(t.grad.cpu().numpy(), g[f"{tag}_grad{n}"]) < RTOL_GRAD, n
    assert cray["rayo"].grad is None and cray["rayd"].grad is None


def _random_sets(n_ray, n_bg, K, seed):
    g = torch.Generator().manual_seed(seed)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).numpy()
    ray = dict(rayo=rnd(n_ray, 3), rayd=rnd(n_ray, 3), zval=rnd(n_ray, 1) + 3, scaling=rnd(n_ray, 3) - 3,
               rotation=rnd(n_ray, 4), opacity=rnd(n_ray, 1, scale=3.0), features_dc=rnd(n_ray, 1, 3),
               features_rest=rnd(n_ray, K - 1, 3)) if n_ray else {}
    bg = dict(xyz=rnd(n_bg, 3, scale=3.0), scaling=rnd(n_bg, 3) - 3, rotation=rnd(n_bg, 4),
              opacity=rnd(n_bg, 1, scale=3.0), features_dc=rnd(n_bg, 1, 3),
              features_rest=rnd(n_bg, K - 1, 3)) if n_bg else {}
    return ray, bg


@pytest.mark.gpu
@pytest.mark.parametrize("n_ray,n_bg,K", [(1, 1, 16), (0, 333, 16), (1000, 0, 16), (5000, 3001, 16), (777, 1234, 9),
                                          (513, 255, 4), (300, 211, 1), (70001, 30003, 16)])
def test_assemble_matches_oracle(n_ray, n_bg, K):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    ray, bg = _random_sets(n_ray, n_bg, K, seed=n_ray + n_bg + K)
    if n_bg > 2:
        bg["opacity"][2] = 40.0                 # saturated sigmoid
    if n_bg == 3001:
        bg["rotation"][1] = 0.0                 # F.normalize's eps clamp: output 0, gradient g / eps
    want, leaves = MO.assemble(ray, bg, torch.float32)
    cray, cbg = _cuda_sets(ray, bg)
    got = _run_fused(cray, cbg)
    ws = [torch.randn(a.shape, generator=torch.Generator().manual_seed(5)) for a in want]
    for name, a, b in zip(ACT, got, want):
        assert a.shape == b.shape and a.is_contiguous()
        assert _rel(a.detach().cpu().numpy(), b.detach().numpy()) < RTOL_ACT, name
    if K > 1:
        # the SH block is a pure copy: bit-exact
        assert torch.equal(got[4].detach().cpu(), want[4].detach())
    sum((a * w).sum() for a, w in zip(want, ws)).backward()
    sum((a * w.cuda()).sum() for a, w in zip(got, ws)).backward()
    for k, t in list(cray.items()) + [("bg_" + k, v) for k, v in cbg.items()]:
        name = k if k.startswith("bg_") else "ray_" + k
        ref = leaves[name].grad
        if ref is None:
            assert t.grad is None, name
            continue
        assert t.grad is not None and t.grad.shape == t.shape, name
        assert _rel(t.grad.cpu().numpy(), ref.numpy()) < RTOL_GRAD, name


class _Cam:
    pass


@pytest.mark.gpu
@pytest.mark.parametrize("split_sh", ["1", "0"])
def test_fused_render_matches_reference_style_render(monkeypatch, split_sh):
    """scgaussian_b200.model.render (fused assembly) against the same operator fed by the reference's chain of torch
    activations (reference gaussian_renderer/__init__.py:55-68 + scene/gaussian_model.py:105-152): same image,
    same gradients on the raw parameters.  split_sh "1" (the default): the operator reads features_dc / features_rest of
    both sets in place and writes dL/dfeatures_* itself (second half); "0": through the assembled
    [P,16,3] copy."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    monkeypatch.setenv("SCGR_SPLIT_SH", split_sh)
    import math
    from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer, model
    from tests.util import assert_grad_close, assert_image_close, make_case
    W, H, n_ray, n_bg = 96, 64, 1500, 700
    case = make_case(n_ray + n_bg, W, H, sh_degree=3, scale_median=0.05, seed=3)
    dev = "cuda"

    def make_pc():
        class PC:
            active_sh_degree = 3
            max_sh_degree = 3
        pc = PC()
        m = case["means3D"]
        g = torch.Generator().manual_seed(9)
        rayo = torch.randn(n_ray, 3, generator=g) * 0.1
        d = m[:n_ray] - rayo
        z = d.norm(dim=1, keepdim=True)
        par = lambda t: torch.nn.Parameter(t.clone().to(dev))
        pc._rayo, pc._rayd, pc._zval = rayo.to(dev), (d / z).to(dev), par(z)
        pc.bg_xyz = par(m[n_ray:])
        sc, ro, op, sh = case["scales"].log(), case["rotations"] * 1.7, torch.logit(case["opacities"]), case["shs"]
        pc._scaling, pc.bg_scaling = par(sc[:n_ray]), par(sc[n_ray:])
        pc._rotation, pc.bg_rotation = par(ro[:n_ray]), par(ro[n_ray:])
        pc._opacity, pc.bg_opacity = par(op[:n_ray]), par(op[n_ray:])
        pc._features_dc, pc.bg_features_dc = par(sh[:n_ray, :1]), par(sh[n_ray:, :1])
        pc._features_rest, pc.bg_features_rest = par(sh[:n_ray, 1:]), par(sh[n_ray:, 1:])
        return pc

    cam = _Cam()
    cam.image_height, cam.image_width = H, W
    cam.FoVx, cam.FoVy = 2 * math.atan(case["tanfovx"]), 2 * math.atan(case["tanfovy"])
    cam.world_view_transform, cam.full_proj_transform = case["viewmatrix"].to(dev), case["projmatrix"].to(dev)
    cam.camera_center = case["campos"].to(dev)
    pipe = _Cam()
    pipe.debug = pipe.compute_cov3D_python = pipe.convert_SHs_python = False
    bgc = torch.tensor([0.1, 0.2, 0.3], device=dev)
    gw = torch.Generator().manual_seed(1)
    w_img, w_d, w_a = (torch.randn(s, generator=gw).to(dev) for s in ((3, H, W), (1, H, W), (1, H, W)))

    def loss_of(out):
        return (out["render"] * w_img).sum() + (out["rendered_depth"] * w_d).sum() + (out["rendered_alpha"] * w_a).sum()

    pc1 = make_pc()
    out1 = model.render(cam, pc1, pipe, bgc)
    loss_of(out1).backward()

    pc2 = make_pc()
    F = torch.nn.functional
    xyz = torch.cat([pc2._rayo + pc2._rayd * pc2._zval, pc2.bg_xyz])
    scal = torch.cat([torch.exp(pc2._scaling), torch.exp(pc2.bg_scaling)])
    rot = torch.cat([F.normalize(pc2._rotation), F.normalize(pc2.bg_rotation)])
    opa = torch.cat([torch.sigmoid(pc2._opacity), torch.sigmoid(pc2.bg_opacity)])
    shs = torch.cat((torch.cat([pc2._features_dc, pc2.bg_features_dc]),
                     torch.cat([pc2._features_rest, pc2.bg_features_rest])), dim=1)
    ssp = torch.zeros_like(xyz, requires_grad=True) + 0
    ssp.retain_grad()
    rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=case["tanfovx"], tanfovy=case["tanfovy"],
                                       bg=bgc, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                       projmatrix=cam.full_proj_transform, sh_degree=3, campos=cam.camera_center,
                                       prefiltered=False, debug=False)
    color, radii, depth, alpha = GaussianRasterizer(raster_settings=rs)(
        means3D=xyz, means2D=ssp, shs=shs, colors_precomp=None, opacities=opa, scales=scal, rotations=rot,
        cov3D_precomp=None)
    out2 = {"render": color, "rendered_depth": depth, "rendered_alpha": alpha}
    loss_of(out2).backward()

    assert set(out1) == {"render", "rendered_depth", "rendered_alpha", "viewspace_points", "visibility_filter", "radii"}
    assert torch.equal(out1["visibility_filter"], out1["radii"] > 0)
    # activations may differ in the last bit between the fused kernel and torch's, which can flip a discrete decision
    # of the rasterizer (alpha < 1/255, T < 1e-4, ceil of the radius): the C oracle, run on the torch-activated
    # inputs, says which pixels / Gaussians sit that close to a threshold -- nothing else may differ (tests/util.py)
    from tests.util import assert_radii_match, flip_sets, run_c_oracle
    act = dict(case, means3D=xyz.detach().cpu(), scales=scal.detach().cpu(), rotations=rot.detach().cpu(),
               opacities=opa.detach().cpu(), shs=shs.detach().cpu(), bg=bgc.cpu())
    flips = flip_sets(run_c_oracle(act)[0])
    assert out1["radii"].dtype == torch.int32
    assert_radii_match("radii", out1["radii"].cpu().numpy(), radii.cpu().numpy(), flips)
    for k in out2:
        assert_image_close(k, out1[k].detach().cpu().numpy(), out2[k].detach().cpu().numpy(), flips)
    assert_grad_close("viewspace_points", out1["viewspace_points"].grad.cpu().numpy(), ssp.grad.cpu().numpy(), flips)
    for n in TRAINED:
        a, b = getattr(pc1, n).grad, getattr(pc2, n).grad
        assert a is not None and b is not None, n
        sl = slice(n_ray, None) if n.startswith("bg_") else slice(0, n_ray)      # the ray-based set comes first
        part = dict(flips, **{k: flips[k][sl] for k in ("gauss_flag", "gauss_margin", "gauss_own")})
        assert_grad_close(n, a.cpu().numpy(), b.cpu().numpy(), part)


@pytest.mark.gpu
def test_adam_replays_reference_golden():
    """The reference's two optimizers (12 groups), 3 steps, one launch per step."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from scgaussian_b200 import _lib, optim
    g = np.load(GOLD)
    params = {n: torch.nn.Parameter(torch.tensor(g[f"adam_p0{n}"], device="cuda")) for n in TRAINED}
    main = optim.Adam([{"params": [params[n]], "lr": 0.0, "name": n} for n in TRAINED if n.startswith("_")],
                      lr=0.0, eps=1e-15)
    bg = optim.Adam([{"params": [params[n]], "lr": 0.0, "name": n} for n in TRAINED if n.startswith("bg_")],
                    lr=0.0, eps=1e-15)
    lib = _lib.load()
    for it in (1, 2, 3):
        for opt in (main, bg):
            for grp in opt.param_groups:
                grp["lr"] = float(g[f"adam_lr{grp['name']}"][it - 1])
        for n in TRAINED:
            params[n].grad = torch.tensor(g[f"adam_g{it}{n}"], device="cuda")
        before = lib.scgr_kernel_launch_count()
        optim.step_all(main, bg)
        assert lib.scgr_kernel_launch_count() - before == 1
        for n in TRAINED:
            want = g[f"adam_p{it}{n}"]
            assert np.abs(params[n].detach().cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max(), (n, it)
    for n in TRAINED:
        st = (main if n.startswith("_") else bg).state[params[n]]
        assert float(st["step"]) == 3.0
        assert _rel(st["exp_avg"].cpu().numpy(), g[f"adam_m3{n}"]) < 1e-6, n
        assert _rel(st["exp_avg_sq"].cpu().numpy(), g[f"adam_v3{n}"]) < 1e-6, n


@pytest.mark.gpu
def test_adam_matches_torch_adam_on_ragged_and_unaligned_groups():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from scgaussian_b200 import optim
    gen = torch.Generator().manual_seed(2)
    sizes = [(1,), (3,), (4096,), (4097,), (100003, 3), (20001, 15, 3), (8191,), (5, 1)]
    lrs = [0.5, 1e-3, 1.6e-4, 5.5e-2, 2e-3, 1e-4, 1.5e-3, 5.5e-3]
    inits = [torch.randn(*s, generator=gen) for s in sizes]
    backing = torch.zeros(9000, device="cuda")

    def make(unaligned_view):
        ps = []
        for k, t in enumerate(inits):
            if unaligned_view and t.numel() == 8191:      # a contiguous view 4 bytes off a 16-byte boundary
                backing[1:8192].copy_(t)
                p = torch.nn.Parameter(backing[1:8192])
            else:
                p = torch.nn.Parameter(t.clone().cuda())
            ps.append(p)
        return ps
    ours_p, ref_p = make(True), make(False)
    assert ours_p[6].data_ptr() % 16 == 4
    ours = optim.Adam([{"params": [p], "lr": lr, "name": str(k)} for k, (p, lr) in enumerate(zip(ours_p, lrs))],
                      lr=0.0, eps=1e-15)
    ref = torch.optim.Adam([{"params": [p], "lr": lr, "name": str(k)} for k, (p, lr) in enumerate(zip(ref_p, lrs))],
                           lr=0.0, eps=1e-15)
    for it in range(6):
        for k, (a, b) in enumerate(zip(ours_p, ref_p)):
            gr = torch.randn(a.shape, generator=gen) * (10.0 ** ((k % 4) - 3))
            if it == 3:
                gr = torch.zeros_like(gr)                  # exp_avg_sq decays, denom -> tiny: eps = 1e-15 matters
            if it == 4 and k == 2:
                a.grad, b.grad = None, None                # torch skips parameters without a gradient
                continue
            a.grad, b.grad = gr.cuda(), gr.cuda()
        ours.step()
        ref.step()
        for k, (a, b) in enumerate(zip(ours_p, ref_p)):
            scale = float(b.detach().abs().max()) + lrs[k]      # parameter scale + the size of one update
            assert float((a.detach() - b.detach()).abs().max()) <= 2e-6 * scale, (it, k)
    assert float(ours.state[ours_p[2]]["step"]) == 5.0 and float(ours.state[ours_p[0]]["step"]) == 6.0
    # state interchange mid-run: torch's state into a fresh fused optimizer, one more step on both
    again = optim.Adam([{"params": [p], "lr": lr, "name": str(k)} for k, (p, lr) in enumerate(zip(ours_p, lrs))],
                       lr=0.0, eps=1e-15)
    for a, b in zip(ours_p, ref_p):
        a.data.copy_(b.data)
    again.load_state_dict(copy.deepcopy(ref.state_dict()))    # load_state_dict itself shares the tensors it is given
    for a, b in zip(ours_p, ref_p):
        gr = torch.randn(a.shape, generator=gen)
        a.grad, b.grad = gr.cuda(), gr.cuda()
    again.step()
    ref.step()
    for k, (a, b) in enumerate(zip(ours_p, ref_p)):
        assert float((a.detach() - b.detach()).abs().max()) <= 2e-6 * (float(b.detach().abs().max()) + lrs[k]), k


class _Stats:
    pass


@pytest.mark.gpu
def test_densification_stats_match_reference_golden_and_oracle():
    """reference train.py:192-193 / scene/gaussian_model.py:932-934 in one launch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from scgaussian_b200 import model
    g = np.load(GOLD)

    def run(grad, radii, accum, denom, maxr, with_filter):
        pc = _Stats()
        pc.xyz_gradient_accum = torch.tensor(accum, device="cuda")
        pc.denom = torch.tensor(denom, device="cuda")
        pc.max_radii2D = torch.tensor(maxr, device="cuda")
        vsp = torch.zeros(grad.shape, device="cuda", requires_grad=True)
        vsp.grad = torch.tensor(grad, device="cuda")
        r = torch.tensor(radii, device="cuda")
        model.add_densification_stats(pc, vsp, (r > 0) if with_filter else None, r)
        return pc.xyz_gradient_accum.cpu().numpy(), pc.denom.cpu().numpy(), pc.max_radii2D.cpu().numpy()

    for with_filter in (False, True):
        a, d, m = run(g["stats_grad"], g["stats_radii"], g["stats_accum0"], g["stats_denom0"], g["stats_maxr0"], with_filter)
        assert np.array_equal(d, g["stats_denom1"]) and np.array_equal(m, g["stats_maxr1"])
        assert _rel(a, g["stats_accum1"]) < 1e-6
    # a large ragged case against the oracle; update_filter only (no radii): max_radii2D untouched
    gen = torch.Generator().manual_seed(8)
    P = 100_003
    grad = (torch.randn(P, 3, generator=gen) * 1e-3).numpy()
    radii = torch.randint(0, 50, (P,), generator=gen).to(torch.int32).numpy()
    radii[::3] = 0
    accum, denom = torch.rand(P, 1, generator=gen).numpy(), torch.randint(0, 9, (P, 1), generator=gen).float().numpy()
    maxr = torch.randint(0, 60, (P,), generator=gen).float().numpy()
    a, d, m = run(grad, radii, accum, denom, maxr, False)
    wa, wd, wm = MO.densification_stats(grad, radii > 0, radii, accum, denom, maxr)
    assert np.array_equal(d, wd) and np.array_equal(m, wm) and _rel(a, wa) < 1e-6
    pc = _Stats()
    pc.xyz_gradient_accum, pc.denom = torch.tensor(accum, device="cuda"), torch.tensor(denom, device="cuda")
    pc.max_radii2D = torch.tensor(maxr, device="cuda")
    vsp = torch.zeros(P, 3, device="cuda", requires_grad=True)
    vsp.grad = torch.tensor(grad, device="cuda")
    filt = torch.tensor(radii > 7, device="cuda")
    model.add_densification_stats(pc, vsp, filt)
    wa, wd, _ = MO.densification_stats(grad, radii > 7, None, accum, denom, maxr)
    assert np.array_equal(pc.denom.cpu().numpy(), wd) and _rel(pc.xyz_gradient_accum.cpu().numpy(), wa) < 1e-6
    assert np.array_equal(pc.max_radii2D.cpu().numpy(), maxr)


@pytest.mark.gpu
def test_prune_points_matches_torch_indexing():
    """scgaussian_b200.densify.prune_points (reference scene/gaussian_model.py:777-820) on the GPU: every per-Gaussian
    array equals `t[valid_mask]` (what the reference evaluates) bit for bit, in 3 launches."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from scgaussian_b200 import _lib, densify, optim
    gen = torch.Generator().manual_seed(13)
    n_ray, n_bg, K = 20011, 9973, 16
    P = n_ray + n_bg

    def rnd(*s):
        return torch.randn(*s, generator=gen).cuda()
    pc = _Stats()
    par = torch.nn.Parameter
    pc._rayo, pc._rayd, pc._zval = rnd(n_ray, 3), rnd(n_ray, 3), par(rnd(n_ray, 1))
    pc._features_dc, pc._features_rest = par(rnd(n_ray, 1, 3)), par(rnd(n_ray, K - 1, 3))
    pc._scaling, pc._rotation, pc._opacity = par(rnd(n_ray, 3)), par(rnd(n_ray, 4)), par(rnd(n_ray, 1))
    pc.bg_xyz, pc.bg_features_dc, pc.bg_features_rest = par(rnd(n_bg, 3)), par(rnd(n_bg, 1, 3)), par(rnd(n_bg, K - 1, 3))
    pc.bg_scaling, pc.bg_rotation, pc.bg_opacity = par(rnd(n_bg, 3)), par(rnd(n_bg, 4)), par(rnd(n_bg, 1))
    inv = {v: k for k, v in densify.GROUP_ATTR.items()}
    main = [a for a in densify.GROUP_ATTR.values() if a.startswith("_")]
    free = [a for a in densify.GROUP_ATTR.values() if a.startswith("bg_")]
    pc.optimizer = optim.Adam([{"params": [getattr(pc, a)], "lr": 1e-3, "name": inv[a]} for a in main], lr=0.0, eps=1e-15)
    pc.optimizer_bg = torch.optim.Adam([{"params": [getattr(pc, a)], "lr": 1e-3, "name": inv[a]} for a in free],
                                       lr=0.0, eps=1e-15)          # torch's optimizer works as well
    for a in main + free:
        getattr(pc, a).grad = rnd(*getattr(pc, a).shape)
    pc.optimizer.step()
    pc.optimizer_bg.step()
    pc.xyz_gradient_accum, pc.denom, pc.max_radii2D = rnd(P, 1), rnd(P, 1).abs(), rnd(P).abs()
    mask = (torch.rand(P, generator=gen) < 0.3).cuda()
    mask[:5] = True
    valid = ~mask
    want = {a: getattr(pc, a).detach()[valid[:n_ray]] for a in main + ["_rayo", "_rayd"]}
    want.update({a: getattr(pc, a).detach()[valid[n_ray:]] for a in free})
    want.update({a: getattr(pc, a)[valid] for a in ("xyz_gradient_accum", "denom", "max_radii2D")})
    want_state = {}
    for opt, names, v in ((pc.optimizer, main, valid[:n_ray]), (pc.optimizer_bg, free, valid[n_ray:])):
        for a in names:
            st = opt.state[getattr(pc, a)]
            want_state[a] = (st["exp_avg"][v], st["exp_avg_sq"][v])
    lib = _lib.load()
    before = lib.scgr_kernel_launch_count()
    densify.prune_points(pc, mask)
    assert lib.scgr_kernel_launch_count() - before == 3
    for a, w in want.items():
        got = getattr(pc, a)
        assert got.shape == w.shape and torch.equal(got.detach(), w), a
    for opt, names in ((pc.optimizer, main), (pc.optimizer_bg, free)):
        for grp in opt.param_groups:
            a = densify.GROUP_ATTR[grp["name"]]
            p = grp["params"][0]
            assert p is getattr(pc, a) and isinstance(p, torch.nn.Parameter) and p.requires_grad
            st = opt.state[p]
            assert torch.equal(st["exp_avg"], want_state[a][0]) and torch.equal(st["exp_avg_sq"], want_state[a][1]), a
            assert float(st["step"]) == 1.0
        assert len(opt.state) == 6
    # the pruned model keeps training: one more step on both optimizers
    for a in main + free:
        getattr(pc, a).grad = torch.ones_like(getattr(pc, a))
    pc.optimizer.step()
    pc.optimizer_bg.step()
    assert float(pc.optimizer.state[pc._zval]["step"]) == 2.0
    # nothing pruned / everything of one set pruned
    densify.prune_points(pc, torch.zeros(pc._zval.shape[0] + pc.bg_xyz.shape[0], dtype=torch.bool, device="cuda"))
    assert pc._zval.shape[0] == int(valid[:n_ray].sum())
    m2 = torch.zeros(pc._zval.shape[0] + pc.bg_xyz.shape[0], dtype=torch.bool, device="cuda")
    m2[pc._zval.shape[0]:] = True
    densify.prune_points(pc, m2)
    assert pc.bg_xyz.shape == (0, 3) and pc.bg_features_rest.shape == (0, K - 1, 3) and pc.max_radii2D.shape[0] == pc._zval.shape[0]
