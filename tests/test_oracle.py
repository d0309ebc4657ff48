"""CPU tests of the oracle itself: it is pinned to every reference-owned vector we have
(tests/golden/reference_anchors.npz, produced by importing the reference's Python helpers) and
its two implementations (torch autograd, C explicit backward) are pinned to each other."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests import util

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_anchors.npz"))


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_matches_reference_eval_sh(deg):
    shs = torch.from_numpy(G["sh_shs"])
    xyz = torch.from_numpy(G["sh_xyz"])
    cam = torch.from_numpy(G["sh_campos"])
    d = xyz - cam[None]
    d = d / d.norm(dim=1, keepdim=True)
    rgb = torch.clamp_min(O.eval_sh_rgb(deg, shs, d) + 0.5, 0.0)
    np.testing.assert_allclose(rgb.numpy(), G[f"sh_rgb_deg{deg}"], rtol=1e-12, atol=1e-12)


def test_projection_matrix_matches_reference():
    fovx, fovy = G["proj_fov"]
    Pm = O.projection_matrix(0.01, 100.0, float(fovx), float(fovy))
    np.testing.assert_allclose(Pm.numpy(), G["proj_matrix"], rtol=1e-6, atol=1e-7)


def test_make_camera_matches_reference_camera():
    R, t = G["xf_R"], G["xf_t"]
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = torch.from_numpy(R.T)
    w2c[:3, 3] = torch.from_numpy(t)
    view = w2c.float().T
    np.testing.assert_allclose(view.numpy(), G["xf_world_view"], rtol=1e-6, atol=1e-6)
    campos = torch.linalg.inv(view)[3, :3]
    np.testing.assert_allclose(campos.numpy(), G["xf_campos"], rtol=1e-5, atol=1e-6)


def test_homogeneous_projection_matches_reference():
    pts = torch.from_numpy(G["xf_points"])
    full = torch.from_numpy(G["xf_full_proj"])
    hom = torch.cat([pts, torch.ones(len(pts), 1)], 1) @ full
    ndc = hom[:, :3] / (hom[:, 3:] + 1e-7)
    np.testing.assert_allclose(ndc.numpy(), G["xf_ndc"], rtol=1e-5, atol=1e-6)
    # and through preprocess(): pixel coords derive from that ndc (A.2)
    W, H = 128, 96
    s = O.Settings(H, W, 0.5, 0.4, torch.zeros(3), 1.0, torch.from_numpy(G["xf_world_view"]), full, 0,
                   torch.from_numpy(G["xf_campos"]))
    P = len(pts)
    geom = O.preprocess(pts, torch.ones(P, 1), s, colors_precomp=torch.ones(P, 3),
                        scales=torch.full((P, 3), 0.01), rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1))
    px = ((torch.from_numpy(G["xf_ndc"])[:, 0] + 1) * W - 1) * 0.5
    np.testing.assert_allclose(geom.means2D[:, 0].numpy(), px.numpy(), rtol=1e-4, atol=1e-3)


def test_covariance_matches_reference_build_scaling_rotation():
    sc = torch.from_numpy(G["cov_scales"])
    rot = torch.from_numpy(G["cov_rots"])
    q = rot / rot.norm(dim=1, keepdim=True)        # the reference normalises inside build_rotation
    np.testing.assert_allclose(O.build_rotation_unnormalized(q).numpy(), G["cov_R"], rtol=1e-5, atol=1e-6)
    S = O.cov3d_from_scale_rot(sc, q, float(G["cov_modifier"]))
    six = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)
    np.testing.assert_allclose(six.numpy(), G["cov_sixvec"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(O.cov3d_from_sixvec(six).numpy(), S.numpy(), rtol=1e-6, atol=1e-12)


# ---------------------------------------------------------------------------------------------
def _torch_run(case, dtype, use_cov=False, use_col=False):
    s = util.oracle_settings(case, dtype)
    leaves = {k: case[k].to(dtype).clone().requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros(case["P"], 3, dtype=dtype, requires_grad=True)
    kw, extra = {}, {}
    if use_cov:
        c3 = O.cov3d_from_scale_rot(case["scales"].to(dtype), case["rotations"].to(dtype), case["scale_modifier"])
        c6 = torch.stack([c3[:, 0, 0], c3[:, 0, 1], c3[:, 0, 2], c3[:, 1, 1], c3[:, 1, 2], c3[:, 2, 2]], -1)
        extra["cov3D_precomp"] = c6.clone().requires_grad_(True)
        kw["cov3D_precomp"] = extra["cov3D_precomp"]
    else:
        kw["scales"], kw["rotations"] = leaves["scales"], leaves["rotations"]
    if use_col:
        g = torch.Generator().manual_seed(7)
        extra["colors_precomp"] = torch.rand(case["P"], 3, generator=g).to(dtype).requires_grad_(True)
        kw["colors_precomp"] = extra["colors_precomp"]
    else:
        kw["shs"] = leaves["shs"]
    out = O.rasterize(leaves["means3D"], m2d, leaves["opacities"], s, return_aux=True, **kw)
    return out, leaves, m2d, extra


@pytest.mark.parametrize("deg,use_cov,use_col,bg,mod", [
    (3, False, False, (0.1, 0.2, 0.3), 1.0),
    (2, True, False, (0.0, 0.0, 0.0), 1.3),
    (1, False, True, (1.0, 1.0, 1.0), 0.7),
    (0, False, False, (0.5, 0.0, 0.2), 1.0),
])
def test_c_oracle_matches_torch_autograd_f64(deg, use_cov, use_col, bg, mod):
    case = util.make_case(250, 72, 56, sh_degree=deg, scale_median=0.06, bg=bg, scale_modifier=mod,
                          w2c=O.yaw_w2c(10.0))
    (color, radii, depth, alpha, aux), leaves, m2d, extra = _torch_run(case, torch.float64, use_cov, use_col)
    gC, gD, gA = O.synth_upstream_grads(case["W"], case["H"], dtype=torch.float64)
    ((color * gC).sum() + (depth * gD).sum() + (alpha * gA).sum()).backward()
    co, (c2, r2, d2, a2), g = util.run_c_oracle(
        case, "f64", grads=(gC, gD, gA),
        colors_precomp=extra["colors_precomp"].detach() if use_col else None,
        cov3D_precomp=extra["cov3D_precomp"].detach() if use_cov else None)
    assert util.rel_err(c2, color.detach()) < 1e-12
    assert util.rel_err(d2, depth.detach()) < 1e-12
    assert util.rel_err(a2, alpha.detach()) < 1e-12
    assert np.array_equal(r2, radii.numpy())
    st = co.state()
    pl, rg = O.tile_lists(aux["geom"], util.oracle_settings(case, torch.float64))
    assert np.array_equal(st["point_list"].astype(np.int64), pl.numpy())
    assert np.array_equal(st["ranges"], rg.numpy())
    assert np.array_equal(st["n_contrib"], aux["n_contrib"].numpy())
    named = {"means3D": leaves["means3D"], "means2D": m2d, "opacities": leaves["opacities"]}
    if use_cov:
        named["cov3D_precomp"] = extra["cov3D_precomp"]
    else:
        named["scales"], named["rotations"] = leaves["scales"], leaves["rotations"]
    named["colors_precomp" if use_col else "shs"] = extra["colors_precomp"] if use_col else leaves["shs"]
    for k, v in named.items():
        # 1e-5: the explicit backward keeps the external rasterizer's 1/(det^2 + 1e-7) (A.10)
        assert util.rel_err(g[k], v.grad.numpy()) < 1e-5, k


def test_gradcheck_fp64_tiny_scene():
    case = util.make_case(6, 20, 18, sh_degree=1, scale_median=0.15, max_sh_degree=1)
    s = util.oracle_settings(case, torch.float64)
    d = torch.float64
    m3 = case["means3D"].to(d).requires_grad_(True)
    op = case["opacities"].to(d).requires_grad_(True)
    sc = case["scales"].to(d).requires_grad_(True)
    ro = case["rotations"].to(d).requires_grad_(True)
    sh = case["shs"].to(d).requires_grad_(True)

    def f(m3, op, sc, ro, sh):
        c, _, dep, a = O.rasterize(m3, None, op, s, shs=sh, scales=sc, rotations=ro)
        return c.sum() * 0.3 + (dep * dep).sum() * 0.1 + a.sum() * 0.2

    assert torch.autograd.gradcheck(f, (m3, op, sc, ro, sh), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0)


def test_render_identities():
    case = util.make_case(300, 64, 48, scale_median=0.05, bg=(0.2, 0.4, 0.6))
    s = util.oracle_settings(case, torch.float64)
    c, r, d, a, aux = O.rasterize(case["means3D"].double(), None, case["opacities"].double(), s,
                                  shs=case["shs"].double(), scales=case["scales"].double(),
                                  rotations=case["rotations"].double(), return_aux=True)
    T = aux["final_T"]
    assert (a[0] - (1 - T)).abs().max() < 1e-12           # alpha == 1 - prod(1 - alpha_i)
    assert (T >= 1e-4 - 1e-15).all()
    # single fully opaque Gaussian at a pixel centre: alpha = min(.99, opacity)
    cam = O.make_camera(32, 32)
    s1 = O.Settings(32, 32, cam["tanfovx"], cam["tanfovy"], torch.zeros(3), 1.0, cam["viewmatrix"],
                    cam["projmatrix"], 0, cam["campos"])
    g1 = O.preprocess(torch.tensor([[0.0, 0.0, 5.0]]), torch.ones(1, 1), s1, colors_precomp=torch.ones(1, 3),
                      scales=torch.full((1, 3), 0.05), rotations=torch.tensor([[1.0, 0, 0, 0]]))
    # mean projects to pixel (15.5, 15.5): between pixel centres
    assert torch.allclose(g1.means2D, torch.tensor([[15.5, 15.5]]), atol=1e-4)


def test_edge_cases_cpu():
    # P = 0 -> zero images, not background filled (section 8b)
    cam = O.make_camera(40, 24)
    s = O.Settings(24, 40, cam["tanfovx"], cam["tanfovy"], torch.ones(3), 1.0, cam["viewmatrix"],
                   cam["projmatrix"], 0, cam["campos"])
    c, r, d, a = O.rasterize(torch.zeros(0, 3), None, torch.zeros(0, 1), s, shs=torch.zeros(0, 1, 3),
                             scales=torch.zeros(0, 3), rotations=torch.zeros(0, 4))
    assert c.abs().sum() == 0 and r.numel() == 0
    # all culled (behind the camera) -> background everywhere, radii 0
    case = util.make_case(50, 40, 24, sh_degree=0, max_sh_degree=0, bg=(0.3, 0.6, 0.9), z_shift=-20.0)
    co, (c2, r2, d2, a2), _ = util.run_c_oracle(case)
    assert (r2 == 0).all() and np.allclose(c2[1], 0.6) and (a2 == 0).all() and co.num_rendered == 0
    # ragged image (W, H not multiples of 16) agrees between the two oracles in fp32
    case = util.make_case(200, 50, 37, scale_median=0.08)
    so = util.oracle_settings(case)
    ct, rt, dt_, at = O.rasterize(case["means3D"], None, case["opacities"], so, shs=case["shs"],
                                  scales=case["scales"], rotations=case["rotations"])
    co, (c2, r2, d2, a2), _ = util.run_c_oracle(case)
    util.assert_image_close("ragged color", c2, ct.numpy(), util.flip_sets(co))
    assert (np.abs(r2 - rt.numpy()) <= 1).all() and (r2 != rt.numpy()).mean() < 0.02
    assert math.isfinite(float(c2.sum()))


def test_f32_f64_oracles_differ_only_where_a_decision_sits_on_its_threshold():
    """Why the parity compares carry a flip allowance at all, and that the allowance is PROVEN per element: the
    float32 and float64 builds of the SAME C oracle differ beyond 1e-3 on a few gradient rows of this scene -- and
    every one of them belongs to a Gaussian that touches a pixel in which a discrete decision (alpha >= 1/255,
    T >= 1e-4) lies inside the uncertainty band of its threshold (oracle/scg_oracle.c: scgo_margins).  Nothing else is excused."""
    case = util.make_case(5000, 378, 504, sh_degree=0, scale_median=0.03, w2c=O.yaw_w2c(5.0))
    grads = O.synth_upstream_grads(378, 504)
    co32, img32, g32 = util.run_c_oracle(case, "f32", grads=grads)
    co64, img64, g64 = util.run_c_oracle(case, "f64", grads=grads)
    flips = util.flip_sets(co32)
    print("flip-prone pixels", flips["pix_flag"].mean(), "flip-affected Gaussians", flips["gauss_flag"].mean(),
          "own", flips["gauss_own"].mean())
    assert 0 < flips["pix_flag"].mean() < 0.05 and 0 < flips["gauss_flag"].mean() < 0.5
    beyond = 0
    for k in ("means3D", "means2D", "opacities", "shs", "scales", "rotations"):
        st = util.assert_grad_close(k, g32[k], g64[k], flips)
        beyond += st["n_beyond_rtol"]
        assert st["max_err_unexcusable"] < 3e-4, (k, st)
    assert beyond > 0          # the scene does contain flips: without the proof this test would fail
    for a, b, name in zip(img32, img64, ("color", "radii", "depth", "alpha")):
        if name != "radii":
            util.assert_image_close(name, a, b, flips)
    util.assert_radii_match("radii", img32[1], img64[1], flips)
