"""BASELINE config 5's loss path -- the match-prior loss on the rendered depth (reference
scene/gaussian_model.py:241-282) and the DTU background term (reference train.py:151-158, :167-168) -- SURVEY.md
section 8(f) row f1 second half + VERDICT r01 item 10.

CPU: the oracle (oracle/prior_oracle.py) against vectors produced by the reference's own method / statements
(tests/golden/prior_golden.npz, made by tests/golden/make_prior_golden.py), and the kernels themselves on the host
(tests/emulation) against the same vectors.  GPU (-m gpu): the fused operators through the C ABI against the vectors,
and the whole config-5 loss -- L1 + D-SSIM + 0.3 match loss + alpha term -- backpropagated through the rasterizer's
depth and alpha outputs against the CPU oracle chain."""
import ctypes as C
import os
import types

import numpy as np
import pytest
import torch

from oracle import prior_oracle as PO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prior_golden.npz"))
NAMES = [str(n) for n in G["match_names"]]
W, H = (int(x) for x in G["match_size"])


@pytest.mark.parametrize("vi", [0, 1, 2])
def test_oracle_matches_reference_golden(vi):
    depth = torch.from_numpy(G[f"match_depth_{vi}"][0]).double().requires_grad_(True)
    loss = PO.match_loss(depth, PO.pairs_from_golden(G, NAMES[vi]), float(W), float(H))
    g, = torch.autograd.grad(loss, depth)
    assert abs(loss.item() - float(G[f"match_loss_{vi}"])) < 2e-7
    ref = G[f"match_grad_{vi}"][0]
    assert np.abs(g.numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
    assert (ref != 0).sum() > 100


@pytest.mark.parametrize("tag", ["a", "b"])
def test_background_oracle_matches_reference_statements(tag):
    mask, gt = PO.dtu_background_mask(G[f"bg_{tag}_gt"])
    assert np.array_equal(mask, G[f"bg_{tag}_mask"]) and mask.sum() > 500
    assert np.array_equal(gt, G[f"bg_{tag}_gt_masked"])
    assert abs(PO.masked_mean(G[f"bg_{tag}_alpha"], mask) - float(G[f"bg_{tag}_alpha_mean"])) < 1e-6


# ---- the kernels on the host (tests/emulation) ----
@pytest.fixture(scope="module")
def emu():
    from tests.emulation import build
    try:
        path = build.build_loss_knn()
    except Exception as e:      # pragma: no cover
        pytest.skip(f"host emulation library not buildable here: {e}")
    lib = C.CDLL(path)
    lib.emu_masked_mean_scratch_bytes.restype = C.c_size_t
    return lib


def _pair_table(name0, dev="cpu"):
    from scgaussian_b200._lib import ScgrMatchPair
    keep, pairs = [], []
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)      # noqa: E731
    for name1 in NAMES:
        if name1 == name0:
            continue
        k = f"view_{name0}_{name1}_"
        arrs = [f(G[k + "uv"]), f(G[k + "rays_o"]), f(G[k + "rays_d"]), f(G[k + "cam_rays_d"]), f(G[f"view_{name1}_{name0}_uv"]),
                f(G[k + "blender_mask"] * G[f"view_{name1}_{name0}_blender_mask"])]
        keep += arrs
        pairs.append(ScgrMatchPair(arrs[0].shape[0], *[a.data_ptr() for a in arrs],
                                   (C.c_float * 12)(*G[f"view_{name1}_w2c"][:3].reshape(-1).tolist()),
                                   (C.c_float * 9)(*G[f"view_{name1}_intr"].reshape(-1).tolist())))
    return (ScgrMatchPair * len(pairs))(*pairs), len(pairs), keep


@pytest.mark.parametrize("vi", [0, 1, 2])
def test_match_loss_kernels_on_host_match_reference_golden(emu, vi):
    table, n, keep = _pair_table(NAMES[vi])
    depth = torch.from_numpy(G[f"match_depth_{vi}"][0]).contiguous()
    scratch, out = torch.zeros(8), torch.full((1,), float("nan"))
    p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
    emu.emu_match_loss_forward(p(depth), H, W, C.c_float(W), C.c_float(H), table, n, p(scratch), p(out))
    assert abs(float(out) - float(G[f"match_loss_{vi}"])) < 1e-6
    grad = torch.full((H, W), float("nan"))
    up = torch.tensor([1.0])
    emu.emu_match_loss_backward(p(depth), H, W, C.c_float(W), C.c_float(H), table, n, p(scratch), p(up), p(grad))
    ref = G[f"match_grad_{vi}"][0]
    assert np.abs(grad.numpy() - ref).max() <= 2e-4 * np.abs(ref).max()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_background_kernels_on_host_match_reference_statements(emu, tag):
    gt = torch.from_numpy(G[f"bg_{tag}_gt"]).contiguous()
    c, h, w = gt.shape
    mask = torch.full((1, h, w), 7, dtype=torch.uint8)
    count = torch.full((1,), float("nan"))
    p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
    emu.emu_bg_mask(p(gt), c, h, w, C.c_float(30.0 / 255.0), 50, p(mask), p(count))
    assert np.array_equal(mask.numpy().astype(bool), G[f"bg_{tag}_mask"])
    assert np.array_equal(gt.numpy(), G[f"bg_{tag}_gt_masked"]) and float(count) == G[f"bg_{tag}_mask"].sum()
    alpha = torch.from_numpy(G[f"bg_{tag}_alpha"]).contiguous()
    n = alpha.numel()
    raw = torch.zeros(int(emu.emu_masked_mean_scratch_bytes(C.c_int64(n))) + 64, dtype=torch.uint8)
    scr = raw[(-raw.data_ptr()) % 64:]
    out2 = torch.full((2,), float("nan"))
    emu.emu_masked_mean_forward(p(alpha), p(mask), C.c_int64(n), p(scr), p(out2))
    assert abs(float(out2[0]) - float(G[f"bg_{tag}_alpha_mean"])) < 1e-6 and float(out2[1]) == float(count)
    g = torch.full((1, h, w), float("nan"))
    emu.emu_masked_mean_backward(p(mask), C.c_int64(n), p(out2), None, p(g))
    assert np.allclose(g.numpy(), G[f"bg_{tag}_alpha_grad"], rtol=1e-6, atol=0)


# ---- GPU: the fused operators through the C ABI ----
def _golden_model(dev):
    """An object with the reference model's `view_gs` bookkeeping (reference scene/gaussian_model.py:356-366), on `dev`."""
    t = lambda a: torch.from_numpy(np.asarray(a)).float().to(dev)      # noqa: E731
    vg = {}
    for n0 in NAMES:
        vg[n0] = {"intr": t(G[f"view_{n0}_intr"]), "w2c": t(G[f"view_{n0}_w2c"]), "width": W, "height": H, "match_infos": {}}
        for n1 in NAMES:
            if n1 != n0:
                k = f"view_{n0}_{n1}_"
                vg[n0]["match_infos"][n1] = {q: t(G[k + q]) for q in ("uv", "rays_o", "rays_d", "cam_rays_d", "blender_mask")}
    return types.SimpleNamespace(view_gs=vg)


@pytest.mark.gpu
@pytest.mark.parametrize("vi", [0, 1, 2])
def test_fused_match_loss_matches_reference_golden(vi):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from scgaussian_b200.losses import get_matchloss_from_renderdepth
    dev = torch.device("cuda:0")
    pc = _golden_model(dev)
    cam0 = types.SimpleNamespace(image_name=NAMES[vi])
    depth = torch.from_numpy(G[f"match_depth_{vi}"]).to(dev).requires_grad_(True)          # [1,H,W] like rendered_depth
    loss = get_matchloss_from_renderdepth(pc, cam0, depth, None)
    (loss * 0.3).backward()
    assert abs(float(loss) - float(G[f"match_loss_{vi}"])) < 1e-6
    ref = 0.3 * G[f"match_grad_{vi}"]
    assert depth.grad.shape == (1, H, W)
    assert np.abs(depth.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_fused_background_term_matches_reference_statements(tag):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from scgaussian_b200.losses import dtu_background_mask, masked_mean
    dev = torch.device("cuda:0")
    gt = torch.from_numpy(G[f"bg_{tag}_gt"]).to(dev).contiguous()
    mask = dtu_background_mask(gt)
    assert mask.dtype == torch.bool and np.array_equal(mask.cpu().numpy(), G[f"bg_{tag}_mask"])
    assert np.array_equal(gt.cpu().numpy(), G[f"bg_{tag}_gt_masked"])
    alpha = torch.from_numpy(G[f"bg_{tag}_alpha"]).to(dev).requires_grad_(True)
    m = masked_mean(alpha, mask)
    m.backward()
    assert abs(float(m) - float(G[f"bg_{tag}_alpha_mean"])) < 1e-6
    assert np.allclose(alpha.grad.cpu().numpy(), G[f"bg_{tag}_alpha_grad"], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_config5_loss_reaches_the_rasterizer_depth_and_alpha_gradients():
    """BASELINE config 5, the actual loss of reference train.py:160-170 on a hybrid-style scene at 96x72: L1 + 0.2 D-SSIM
    on the image, + 0.3 x the match-prior loss on the RENDERED DEPTH, + the mean of the RENDERED ALPHA over the DTU
    background mask -- all fused operators of this repo -- backpropagated through the rasterizer.  The parameter
    gradients must equal the CPU oracle's rasterizer backward fed with the upstream gradients of the oracle loss chain
    (the reference-pinned restatements of oracle/loss_oracle.py + oracle/prior_oracle.py) evaluated on the oracle's own
    images: the depth and alpha gradient paths are exercised end to end."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import loss_oracle as LO
    from oracle import torch_oracle as O
    from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer
    from scgaussian_b200.losses import dtu_background_mask, get_matchloss_from_renderdepth, masked_mean, photometric_loss
    from tests import util
    dev = torch.device("cuda:0")
    P = 2500
    case = util.make_case(P, W, H, sh_degree=2, max_sh_degree=3, scale_median=0.09, bg=(0.0, 0.0, 0.0), seed=17, z_shift=2.0)
    gen = torch.Generator().manual_seed(5)
    gt = torch.rand(3, H, W, generator=gen) * 0.6
    gt[:, :, : W // 4] *= 0.1                                   # a dark band: the DTU background
    leaves = {k: case[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    s = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=case["tanfovx"], tanfovy=case["tanfovy"],
                                      bg=case["bg"].to(dev), scale_modifier=1.0, viewmatrix=case["viewmatrix"].to(dev),
                                      projmatrix=case["projmatrix"].to(dev), sh_degree=2, campos=case["campos"].to(dev),
                                      prefiltered=False, debug=False)
    color, radii, depth, alpha = GaussianRasterizer(s)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                      shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    gt_dev = gt.to(dev).contiguous()
    bg_mask = dtu_background_mask(gt_dev)                         # zeroes gt_dev where masked (train.py:158)
    pc = _golden_model(dev)
    cam0 = types.SimpleNamespace(image_name=NAMES[1])
    loss = photometric_loss(color, gt_dev, 0.2) + 0.3 * get_matchloss_from_renderdepth(pc, cam0, depth, None) \
        + masked_mean(alpha, bg_mask)
    loss.backward()
    torch.cuda.synchronize()
    # ---- the oracle chain, float64 losses on the oracle's float32 images ----
    co, (c2, r2, d2, a2), _ = util.run_c_oracle(case, "f32")
    mask_o, gt_o = PO.dtu_background_mask(gt.numpy())
    assert np.array_equal(bg_mask.cpu().numpy(), mask_o) and mask_o.sum() > 200
    ci = torch.from_numpy(c2).double().requires_grad_(True)
    di = torch.from_numpy(d2).double().requires_grad_(True)
    ai = torch.from_numpy(a2).double().requires_grad_(True)
    lo = LO.photometric_loss(ci, torch.from_numpy(gt_o).double(), 0.2) \
        + 0.3 * PO.match_loss(di[0], PO.pairs_from_golden(G, NAMES[1]), float(W), float(H)) \
        + ai[torch.from_numpy(mask_o)].mean()
    gc, gd, ga = torch.autograd.grad(lo, (ci, di, ai))
    assert abs(float(loss) - float(lo)) < 5e-5 * abs(float(lo))
    assert float(gd.abs().max()) > 0 and float(ga.abs().max()) > 0
    g2 = co.backward(gc.float().numpy(), gd.float().numpy(), ga.float().numpy())
    flips = util.flip_sets(co)
    util.assert_radii_match("radii", radii.cpu().numpy(), r2, flips)
    for k, v in leaves.items():
        util.assert_grad_close(k, v.grad.cpu().numpy(), g2[k].reshape(tuple(v.shape)), flips, normwise=True)
    util.assert_grad_close("means2D", m2d.grad.cpu().numpy(), g2["means2D"], flips, normwise=True)
