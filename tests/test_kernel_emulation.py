"""The kernels of libscgr.so executed on the HOST, thread for thread: scgaussian_b200/csrc/{preprocess,binning,render,
loss,model,knn}.cu compiled with g++ through a small CUDA shim (tests/emulation/: one real thread per CUDA thread, real
block barriers, warp votes / shuffles on per-warp barriers, shared memory as statics; inline PTX replaced statement by
statement, the TMA plumbing by a synchronous copy) and compared with the same oracles, reference-generated golden
vectors and tolerances as the GPU tests -- up to the whole operator, forward and backward, against the C oracle.

TEST INFRASTRUCTURE, CPU suite only: it catches indexing / bounds / protocol / staging mistakes before a GPU is
available.  It is NOT a CPU path of the product (nothing under scgaussian_b200/ can reach it) and it proves nothing
about performance or about what only the hardware does -- the `-m gpu` tests remain the parity gate."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from scgaussian_b200 import _lib as L
from tests import test_model as TM

GOLD = TM.GOLD


@pytest.fixture(scope="module")
def emu():
    from tests.emulation import build
    try:
        path = build.build()
    except Exception as e:      # pragma: no cover  (no g++ / no C++20 <barrier>)
        pytest.skip(f"host emulation library not buildable here: {e}")
    return C.CDLL(path)


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _model(ray, bg, K):
    tr = {k: torch.tensor(np.asarray(v), dtype=torch.float32) for k, v in ray.items()}
    tb = {k: torch.tensor(np.asarray(v), dtype=torch.float32) for k, v in bg.items()}

    def mk(d):
        if not d:
            return L.ScgrModelSet(0, *[None] * 9)
        return L.ScgrModelSet(int(d["scaling"].shape[0]), _p(d.get("xyz")), _p(d.get("rayo")), _p(d.get("rayd")),
                              _p(d.get("zval")), _p(d["scaling"]), _p(d["rotation"]), _p(d["opacity"]),
                              _p(d["features_dc"]), _p(d.get("features_rest")))
    return L.ScgrModel(K - 1, (L.ScgrModelSet * 2)(mk(tr), mk(tb))), (tr, tb)


# staged SH copy: K = 16, 4; flat copy: K = 9, 1; chunk boundaries of the staged copy: 63 / 64 / 65 / 128 Gaussians
@pytest.mark.parametrize("n_ray,n_bg,K", [(1, 1, 16), (0, 133, 16), (200, 0, 16), (321, 190, 16), (130, 65, 4),
                                          (64, 128, 16), (63, 0, 16), (77, 123, 9), (30, 21, 1)])
def test_assemble_kernels_on_host_match_oracle(emu, n_ray, n_bg, K):
    ray, bg = TM._random_sets(n_ray, n_bg, K, seed=n_ray + n_bg + K)
    if n_bg > 2:
        bg["opacity"][2] = 40.0
        bg["rotation"][1] = 0.0
    want, leaves = MO.assemble(ray, bg, torch.float32)
    model, keep = _model(ray, bg, K)
    P = n_ray + n_bg
    outs = [torch.full(s, float("nan")) for s in ((P, 3), (P, 3), (P, 4), (P, 1), (P, K, 3))]
    out = L.ScgrActivated(*[t.data_ptr() for t in outs])
    emu.emu_assemble_forward(C.byref(model), C.byref(out))
    for name, a, b in zip(TM.ACT, outs, want):
        assert not torch.isnan(a).any(), name
        assert TM._rel(a.numpy(), b.detach().numpy()) < TM.RTOL_ACT, name
    assert torch.equal(outs[4], want[4].detach())                       # the SH block is a pure copy
    ws = [torch.randn(a.shape, generator=torch.Generator().manual_seed(5)) for a in want]
    sum((a * w).sum() for a, w in zip(want, ws)).backward()
    gin = L.ScgrActivatedGrads(*[w.data_ptr() for w in ws])

    def grads(n, is_ray):
        nan = lambda *s: torch.full(s, float("nan"))     # noqa: E731
        return dict(xyz=None if is_ray else nan(n, 3), zval=nan(n, 1) if is_ray else None, scaling=nan(n, 3),
                    rotation=nan(n, 4), opacity=nan(n, 1), features_dc=nan(n, 1, 3), features_rest=nan(n, K - 1, 3))
    d_ray, d_bg = grads(n_ray, True), grads(n_bg, False)

    def gset(d, n):
        if not n:
            return L.ScgrModelSetGrads(*[None] * 7)
        return L.ScgrModelSetGrads(_p(d["xyz"]), _p(d["zval"]), _p(d["scaling"]), _p(d["rotation"]), _p(d["opacity"]),
                                   _p(d["features_dc"]), _p(d["features_rest"]))
    og = L.ScgrModelGrads((L.ScgrModelSetGrads * 2)(gset(d_ray, n_ray), gset(d_bg, n_bg)))
    emu.emu_assemble_backward(C.byref(model), C.byref(gin), C.byref(og))
    for prefix, d, n in (("ray_", d_ray, n_ray), ("bg_", d_bg, n_bg)):
        for k, t in d.items():
            if t is None or not n:
                continue
            assert not torch.isnan(t).any(), prefix + k                  # every output element written
            assert TM._rel(t.numpy(), leaves[prefix + k].grad.numpy()) < TM.RTOL_GRAD, prefix + k


def test_adam_kernel_on_host_replays_reference_golden(emu):
    g = np.load(GOLD)
    ps = {n: torch.tensor(g[f"adam_p0{n}"]) for n in TM.TRAINED}
    ms = {n: torch.zeros_like(ps[n]) for n in ps}
    vs = {n: torch.zeros_like(ps[n]) for n in ps}
    for it in (1, 2, 3):
        gr = {n: torch.tensor(g[f"adam_g{it}{n}"]) for n in ps}
        tab = (L.ScgrAdamGroup * 12)(*[L.ScgrAdamGroup(ps[n].data_ptr(), gr[n].data_ptr(), ms[n].data_ptr(), vs[n].data_ptr(),
                                                       ps[n].numel(), float(g[f"adam_lr{n}"][it - 1]), it) for n in TM.TRAINED])
        emu.emu_adam_step(tab, 12, C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        for n in ps:
            want = g[f"adam_p{it}{n}"]
            assert np.abs(ps[n].numpy() - want).max() <= 1e-6 * np.abs(want).max(), (n, it)
    for n in ps:
        assert TM._rel(ms[n].numpy(), g[f"adam_m3{n}"]) < 1e-6 and TM._rel(vs[n].numpy(), g[f"adam_v3{n}"]) < 1e-6, n


def test_adam_kernel_on_host_ragged_and_unaligned(emu):
    gen = torch.Generator().manual_seed(2)
    sizes = [(1,), (3,), (4096,), (4097,), (9001, 3), (8191,)]
    lrs = [0.5, 1e-3, 1.6e-4, 5.5e-2, 2e-3, 1.5e-3]
    inits = [torch.randn(*s, generator=gen) for s in sizes]
    backing = torch.zeros(9000)
    ours = [t.clone() for t in inits]
    backing[1:8192].copy_(inits[5])
    ours[5] = backing[1:8192]                       # 4 bytes off 16-byte alignment: the scalar path
    assert ours[5].data_ptr() % 16 == 4
    ref_p = [torch.nn.Parameter(t.clone()) for t in inits]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref_p, lrs)], lr=0.0, eps=1e-15)
    M, V = [torch.zeros_like(t) for t in inits], [torch.zeros_like(t) for t in inits]
    for it in range(1, 4):
        grs = [torch.randn(t.shape, generator=gen) * 10.0 ** ((k % 4) - 3) for k, t in enumerate(inits)]
        for p, gr in zip(ref_p, grs):
            p.grad = gr.clone()
        ref.step()
        tab = (L.ScgrAdamGroup * len(ours))(*[L.ScgrAdamGroup(ours[k].data_ptr(), grs[k].data_ptr(), M[k].data_ptr(),
                                                              V[k].data_ptr(), ours[k].numel(), lrs[k], it)
                                              for k in range(len(ours))])
        emu.emu_adam_step(tab, len(ours), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        for k in range(len(ours)):
            scale = float(ref_p[k].detach().abs().max()) + lrs[k]
            assert float((ours[k] - ref_p[k].detach()).abs().max()) <= 2e-6 * scale, (it, k)
    assert float(backing[0]) == 0.0 and float(backing[8192:].abs().max()) == 0.0      # nothing written outside the view


def test_densification_stats_kernel_on_host_matches_reference_golden(emu):
    g = np.load(GOLD)
    a, d, m = (torch.tensor(g[k]) for k in ("stats_accum0", "stats_denom0", "stats_maxr0"))
    gr, r = torch.tensor(g["stats_grad"]), torch.tensor(g["stats_radii"])
    emu.emu_densification_stats(_p(gr), None, _p(r), len(r), _p(a), _p(d), _p(m))
    assert np.array_equal(d.numpy(), g["stats_denom1"]) and np.array_equal(m.numpy(), g["stats_maxr1"])
    assert TM._rel(a.numpy(), g["stats_accum1"]) < 1e-6
    # explicit filter, no radii: max_radii2D untouched
    a, d = torch.tensor(g["stats_accum0"]), torch.tensor(g["stats_denom0"])
    f = r > 7
    emu.emu_densification_stats(_p(gr), _p(f), None, len(r), _p(a), _p(d), None)
    wa, wd, _ = MO.densification_stats(g["stats_grad"], f.numpy(), None, g["stats_accum0"], g["stats_denom0"], g["stats_maxr0"])
    assert np.array_equal(d.numpy(), wd) and TM._rel(a.numpy(), wa) < 1e-6


@pytest.mark.parametrize("n_in,keep", [(1, 1.0), (700, 0.5), (9001, 0.9), (500, 0.0), (3333, 0.01)])
def test_gather_rows_kernel_on_host_matches_torch_indexing(emu, n_in, keep):
    gen = torch.Generator().manual_seed(n_in)
    mask = torch.rand(n_in, generator=gen) < keep
    idx = mask.nonzero().squeeze(1)
    shapes = [(1,), (3,), (1, 3), (15, 3), (4,), (), (300,), (257,)]
    srcs = [torch.randn(n_in, *s, generator=gen) for s in shapes]
    dsts = [torch.full((len(idx), *s), float("nan")) for s in shapes]
    if len(idx):
        tab = (L.ScgrRowGather * len(srcs))(*[L.ScgrRowGather(a.data_ptr(), b.data_ptr(), a.numel() // n_in)
                                              for a, b in zip(srcs, dsts)])
        emu.emu_gather_rows(tab, len(srcs), _p(idx), C.c_int64(len(idx)))
    for a, b in zip(srcs, dsts):
        assert torch.equal(b, a[mask])


def test_copy_segments_kernel_on_host(emu):
    gen = torch.Generator().manual_seed(0)
    big = torch.full((40000,), float("nan"))
    segs, off = [], 1
    for k, n in enumerate([1, 3, 4095, 4096, 4097, 12289, 0, 8192]):
        src = torch.randn(n, generator=gen) if k % 3 != 2 else None            # every third segment is a zero fill
        if k % 2 == 0:
            off += (4 - off % 4) % 4                                             # 16-byte aligned destination
        segs.append((off, n, src))
        off += n + 3
    tab = (L.ScgrSegmentCopy * len(segs))(*[L.ScgrSegmentCopy(big[o:o + n].data_ptr() if n else None,
                                                               None if s is None else s.data_ptr(), n) for o, n, s in segs])
    emu.emu_copy_segments(tab, len(segs))
    covered = torch.zeros(40000, dtype=torch.bool)
    for o, n, s in segs:
        assert torch.equal(big[o:o + n], s if s is not None else torch.zeros(n)), (o, n)
        covered[o:o + n] = True
    assert torch.isnan(big[~covered]).all()                                      # nothing written outside the segments
    assert {o % 4 for o, n, s in segs if n} > {0}                                # both the float4 and the scalar path ran


# ------------------------------------------------------------------------------------------------------------------
# scgaussian_b200/csrc/preprocess.cu on the host: the per-Gaussian half of the rasterizer (SURVEY.md section 8a rows
# a9 / a16, Appendix A.1-A.5 and A.10) against the torch oracle's differentiable `preprocess`
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_pre():
    from tests.emulation import build
    try:
        path = build.build_preprocess()
    except Exception as e:      # pragma: no cover
        pytest.skip(f"host emulation library not buildable here: {e}")
    lib = C.CDLL(path)
    lib.emu_geometry_bytes.restype = C.c_size_t
    return lib


def _host_scene(P, W, H, sh_degree, seed, max_sh_degree=3, **kw):
    from tests import util
    case = util.make_case(P, W, H, sh_degree=sh_degree, max_sh_degree=max_sh_degree, seed=seed, **kw)
    t = {k: case[k].contiguous() for k in ("means3D", "opacities", "shs", "scales", "rotations", "bg", "viewmatrix",
                                           "projmatrix", "campos")}
    view = L.ScgrView(H, W, case["tanfovx"], case["tanfovy"], t["bg"].data_ptr(), case["scale_modifier"],
                      t["viewmatrix"].data_ptr(), t["projmatrix"].data_ptr(), sh_degree, t["campos"].data_ptr(), 0, 0)
    g = L.ScgrGaussians(P, int(t["shs"].shape[1]), t["means3D"].data_ptr(), t["opacities"].data_ptr(), t["shs"].data_ptr(),
                        None, t["scales"].data_ptr(), t["rotations"].data_ptr(), None)
    return case, t, view, g


def _geometry(emu_pre, P):
    geom = torch.zeros(int(emu_pre.emu_geometry_bytes(P)) + 64, dtype=torch.uint8)
    base = (-geom.data_ptr()) % 64                      # 64-byte aligned start inside the tensor
    off = (C.c_size_t * 8)()
    emu_pre.emu_geometry_offsets(P, off)

    def view(k, nbytes, dtype):
        return geom[base + off[k]: base + off[k] + nbytes].view(dtype)
    return geom, geom.data_ptr() + base, view


@pytest.mark.parametrize("P,W,H,deg,max_deg,smed,yaw", [(1500, 203, 149, 3, 3, 0.04, 8.0), (700, 64, 48, 1, 1, 0.08, 0.0),
                                                        (900, 160, 120, 0, 2, 0.05, -5.0)])
def test_preprocess_forward_on_host_matches_oracle(emu_pre, P, W, H, deg, max_deg, smed, yaw):
    from oracle import torch_oracle as O
    from tests import util
    case, t, view, g = _host_scene(P, W, H, deg, seed=P, max_sh_degree=max_deg, scale_median=smed, w2c=O.yaw_w2c(yaw),
                                   z_shift=-1.9)          # some Gaussians behind the near plane
    geom, gptr, gv = _geometry(emu_pre, P)
    radii = torch.full((P,), -7, dtype=torch.int32)
    emu_pre.emu_preprocess_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(radii.data_ptr()))
    want = O.preprocess(t["means3D"], t["opacities"], util.oracle_settings(case), shs=t["shs"], scales=t["scales"],
                        rotations=t["rotations"])
    rec = gv(0, P * 48, torch.float32).view(P, 12).numpy()
    r2 = want.radii.numpy()
    n_rad = int((radii.numpy() != r2).sum())
    assert n_rad <= max(2, P // 1000) and np.abs(radii.numpy() - r2).max() <= 1, n_rad      # ceil() of a last-bit difference
    both = (r2 > 0) & (radii.numpy() > 0)
    assert both.sum() > P // 3 and (r2 == 0).sum() > 0
    log2e = 1.4426950408889634
    conic = want.conic.numpy()
    assert util.rel_err(rec[both][:, 0:2], want.means2D.numpy()[both]) < 1e-5
    assert util.rel_err(rec[both][:, 2] / (-0.5 * log2e), conic[both][:, 0]) < 1e-4
    assert util.rel_err(rec[both][:, 3] / -log2e, conic[both][:, 1]) < 1e-4
    assert util.rel_err(rec[both][:, 4] / (-0.5 * log2e), conic[both][:, 2]) < 1e-4
    assert np.array_equal(rec[both][:, 5], t["opacities"].numpy()[both, 0])
    assert util.rel_err(rec[both][:, 6], want.depth.numpy()[both]) < 1e-6
    assert util.rel_err(rec[both][:, 8:11], want.rgb.numpy()[both]) < 1e-5
    bits = rec[:, 11].copy().view(np.uint32)
    assert np.array_equal((bits & 0x0FFFFFFF).astype(np.int32)[both], radii.numpy()[both])
    flags = (bits >> 28)[both]
    clamped = want.clamped.numpy()[both]
    agree = ((flags & 1) > 0) == clamped[:, 0]
    assert agree.mean() > 0.995                          # a colour within rounding of 0 may clamp on one side only
    # culling only removes tiles: never more than the reference's rectangle, and the mask agrees with the count
    touched = gv(1, P * 4, torch.int32).numpy()
    assert (touched[radii.numpy() == r2] <= want.tiles_touched.numpy()[radii.numpy() == r2]).all() and touched.sum() > 0
    assert (touched[radii.numpy() == 0] == 0).all()
    rect = gv(2, P * 8, torch.int32).view(P, 2).numpy().astype(np.int64)
    x0, y0, x1, y1 = rect[:, 0] & 0xFFFF, rect[:, 0] >> 16, rect[:, 1] & 0xFFFF, rect[:, 1] >> 16
    same = radii.numpy() == r2
    assert np.array_equal(np.stack([x0, y0], 1)[both & same], want.rect_min.numpy()[both & same])
    assert np.array_equal(np.stack([x1, y1], 1)[both & same], want.rect_max.numpy()[both & same])
    mask = gv(3, P * 8, torch.int64).numpy()
    small = ((x1 - x0) * (y1 - y0) <= 64) & (radii.numpy() > 0)
    popc = np.array([bin(int(m) & (2 ** 64 - 1)).count("1") for m in mask])
    assert np.array_equal(popc[small], touched[small])

    # depth keys + the digit histograms of the four radix passes (the other kernel that must agree bit for bit)
    emu_pre.emu_depth_keys(C.byref(view), C.byref(g), C.c_void_p(gptr))
    keys = gv(4, P * 4, torch.int32).numpy().view(np.uint32)
    front = rec[:, 6].view(np.uint32)
    vis = radii.numpy() > 0
    assert np.array_equal(keys[vis], front[vis])         # record depth and sort key: identical bits
    zview = (t["means3D"].double() @ t["viewmatrix"].double()[:3, 2] + t["viewmatrix"].double()[3, 2]).numpy()
    assert (keys[zview <= 0.2 - 1e-6] == 0xFFFFFFFF).all()
    assert np.array_equal(gv(5, P * 4, torch.int32).numpy(), np.arange(P))
    words = 260 + 256 * ((P + 4095) // 4096)
    sweep = gv(7, 4 * words * 4, torch.int32).view(4, words).numpy()
    for p in range(4):
        assert np.array_equal(sweep[p, :256], np.bincount((keys >> (8 * p)) & 255, minlength=256))

    # markVisible (A.1 only)
    present = torch.zeros(P, dtype=torch.uint8)
    emu_pre.emu_mark_visible(C.c_void_p(t["means3D"].data_ptr()), P, C.c_void_p(t["viewmatrix"].data_ptr()),
                             C.c_void_p(present.data_ptr()))
    assert np.array_equal(present.numpy().astype(bool), keys != 0xFFFFFFFF)


@pytest.mark.parametrize("P,W,H,deg,max_deg,smed", [(1100, 203, 149, 3, 3, 0.04), (500, 64, 48, 1, 1, 0.08), (640, 120, 90, 2, 3, 0.05)])
def test_preprocess_backward_on_host_matches_autograd_of_the_oracle(emu_pre, P, W, H, deg, max_deg, smed):
    """A.10 in isolation.  The kernel's input is the per-Gaussian accumulator render-backward leaves behind (raw sums
    over pixel pairs, scgaussian_b200/csrc/common.cuh `ScreenGrad`); for random accumulators its outputs must equal the
    autograd gradients of the (fp64) oracle's `preprocess` under the linear functional those sums stand for:
        dL/dmean2D = -(A Sx + B Sy, C Sy + B Sx) px,  dL/dconic = (-SA/2, -SB, -SC/2),  dL/dopacity = Su / opacity."""
    from oracle import torch_oracle as O
    from tests import util
    case, t, view, g = _host_scene(P, W, H, deg, seed=P + 1, max_sh_degree=max_deg, scale_median=smed, w2c=O.yaw_w2c(6.0),
                                   z_shift=-1.9)
    # push some Gaussians beyond 1.3 tan(fov) sideways and make them large enough to still reach the image: the A.4
    # clamp is active for them (A.10: their t.x / t.y are constants in the backward)
    t["means3D"][5:60, 0] *= 1.6
    t["means3D"][60:90, 1] *= 1.7
    t["scales"][5:90] *= 12.0
    geom, gptr, gv = _geometry(emu_pre, P)
    radii = torch.zeros(P, dtype=torch.int32)
    emu_pre.emu_preprocess_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(radii.data_ptr()))
    gen = torch.Generator().manual_seed(3)
    acc = torch.randn(P, 12, generator=gen, dtype=torch.float32) * torch.tensor(
        [1e-2, 1e-2, 1e-1, 1e-1, 1e-1, 1e-2, 1e-3, 0, 1e-2, 1e-2, 1e-2, 0])
    acc[::7] = 0.0                                   # Gaussians nothing blended: the kernel skips them and writes zeros
    acc[radii == 0] = 0.0
    gv(6, P * 48, torch.float32).view(P, 12).copy_(acc)
    M = int(t["shs"].shape[1])
    outs = {"means3D": torch.full((P, 3), float("nan")), "means2D": torch.full((P, 3), float("nan")),
            "shs": torch.full((P, M, 3), float("nan")), "opacities": torch.full((P, 1), float("nan")),
            "scales": torch.full((P, 3), float("nan")), "rotations": torch.full((P, 4), float("nan"))}
    grads = L.ScgrGrads(outs["means3D"].data_ptr(), outs["means2D"].data_ptr(), outs["shs"].data_ptr(), None,
                        outs["opacities"].data_ptr(), outs["scales"].data_ptr(), outs["rotations"].data_ptr(), None)
    emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.byref(grads))
    for k, v in outs.items():
        assert not torch.isnan(v).any(), k            # every gradient tensor is written in full

    leaves = {k: t[k].double().clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    geo = O.preprocess(leaves["means3D"], leaves["opacities"], util.oracle_settings(case, torch.float64), shs=leaves["shs"],
                       scales=leaves["scales"], rotations=leaves["rotations"])
    vis = torch.from_numpy(radii.numpy() > 0) & geo.visible
    pv = torch.cat([t["means3D"].double(), torch.ones(P, 1, dtype=torch.float64)], 1) @ t["viewmatrix"].double()
    clamp_active = ((pv[:, 0] / pv[:, 2]).abs() > 1.3 * case["tanfovx"]) | ((pv[:, 1] / pv[:, 2]).abs() > 1.3 * case["tanfovy"])
    assert int((clamp_active & vis & (acc.abs().sum(1) > 0)).sum()) >= 10
    a = acc.double()
    A, B, Cc = (geo.conic[:, k].detach() for k in range(3))
    gmx, gmy = -(A * a[:, 0] + B * a[:, 1]), -(Cc * a[:, 1] + B * a[:, 0])
    per = (gmx * geo.means2D[:, 0] + gmy * geo.means2D[:, 1]
           - 0.5 * a[:, 2] * geo.conic[:, 0] - a[:, 3] * geo.conic[:, 1] - 0.5 * a[:, 4] * geo.conic[:, 2]
           + a[:, 6] * geo.depth + (a[:, 8:11] * geo.rgb).sum(1)
           + a[:, 5] / leaves["opacities"][:, 0].detach() * leaves["opacities"][:, 0])
    torch.where(vis, per, torch.zeros_like(per)).sum().backward()
    for k in ("means3D", "opacities", "shs", "scales", "rotations"):
        util.assert_grad_close(k, outs[k].numpy(), leaves[k].grad.numpy())
    want2d = torch.stack([gmx * 0.5 * W, gmy * 0.5 * H, torch.zeros_like(gmx)], 1) * vis[:, None]
    util.assert_grad_close("means2D", outs["means2D"].numpy(), want2d.numpy())
    assert float(outs["shs"][:, (deg + 1) ** 2:].abs().max() if M > (deg + 1) ** 2 else 0.0) == 0.0   # rows beyond the active degree
    skipped = (acc.abs().sum(1) == 0).numpy()
    assert skipped.sum() > P // 8 and float(outs["means3D"][torch.from_numpy(skipped)].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------------------------
# scgaussian_b200/csrc/loss.cu and knn.cu on the host
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_loss():
    from tests.emulation import build
    try:
        path = build.build_loss_knn()
    except Exception as e:      # pragma: no cover
        pytest.skip(f"host emulation library not buildable here: {e}")
    lib = C.CDLL(path)
    lib.emu_photometric_scratch_bytes.restype = C.c_size_t
    return lib


def _host_loss(emu_loss, img, gt, lam, upstream=None):
    """(Ll1, ssim, loss, dL/dimage) of the photometric kernels run on the host; upstream: device-scalar stand-in."""
    x = torch.tensor(np.asarray(img), dtype=torch.float32).contiguous()
    y = torch.tensor(np.asarray(gt), dtype=torch.float32).contiguous()
    c, h, w = x.shape
    scratch = torch.zeros(int(emu_loss.emu_photometric_scratch_bytes(c, h, w)) + 256, dtype=torch.uint8)
    sptr = scratch.data_ptr() + (-scratch.data_ptr()) % 256
    out3 = torch.full((3,), float("nan"))
    emu_loss.emu_photometric_forward(_p(x), _p(y), c, h, w, C.c_float(lam), C.c_void_p(sptr), 1, _p(out3))
    grad = torch.full_like(x, float("nan"))
    up = None if upstream is None else torch.tensor([upstream], dtype=torch.float32)
    emu_loss.emu_photometric_backward(_p(x), _p(y), c, h, w, C.c_float(lam), C.c_void_p(sptr), _p(up), _p(grad))
    return float(out3[0]), float(out3[1]), float(out3[2]), grad.numpy()


# SCGR_LOSS_VARIANT: 0 = 32x32 tiles, 1 = streaming column strips, the default (scgaussian_b200/csrc/loss.cu reads it on every launch)
LOSS_VARIANTS = ["0", "1"]


@pytest.mark.parametrize("variant", LOSS_VARIANTS)
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_photometric_kernels_on_host_match_reference_golden(emu_loss, name, variant, monkeypatch):
    """reference train.py:160-161 with utils/loss_utils.py l1_loss / ssim: the vectors in tests/golden/loss_golden.npz were
    produced by the reference's own functions + autograd."""
    from tests import test_loss as TL
    monkeypatch.setenv("SCGR_LOSS_VARIANT", variant)
    g = np.load(TL.GOLD)
    ll1, s, loss, g_loss = _host_loss(emu_loss, g[f"{name}_img"], g[f"{name}_gt"], 0.2)
    assert abs(ll1 - float(g[f"{name}_l1"])) < TL.TOL_VAL
    assert abs(s - float(g[f"{name}_ssim"])) < TL.TOL_VAL
    assert abs(loss - float(g[f"{name}_loss"])) < TL.TOL_VAL
    assert TL._rel(g_loss, g[f"{name}_g_loss"]) < TL.RTOL_GRAD
    # the two parts through the (lambda, upstream) pairs the host side uses: ssim alone = lambda 1, upstream -1
    _, _, _, g_ssim = _host_loss(emu_loss, g[f"{name}_img"], g[f"{name}_gt"], 1.0, upstream=-1.0)
    _, _, _, g_l1 = _host_loss(emu_loss, g[f"{name}_img"], g[f"{name}_gt"], 0.0)
    assert TL._rel(g_ssim, g[f"{name}_g_ssim"]) < TL.RTOL_GRAD and TL._rel(g_l1, g[f"{name}_g_l1"]) < 1e-6


@pytest.mark.parametrize("variant", LOSS_VARIANTS)
@pytest.mark.parametrize("shape", [(3, 1, 1), (3, 5, 7), (1, 16, 16), (3, 17, 33), (6, 40, 24), (2, 70, 131), (1, 34, 117),
                                   (2, 45, 116), (1, 100, 252)])
def test_photometric_kernels_on_host_match_oracle_f64(emu_loss, shape, variant, monkeypatch):
    from tests import test_loss as TL
    monkeypatch.setenv("SCGR_LOSS_VARIANT", variant)
    gen = torch.Generator().manual_seed(sum(shape))
    gt = torch.rand(*shape, generator=gen)
    img = (gt + 0.2 * torch.randn(*shape, generator=gen)).clamp(0, 1)
    o = TL._oracle_all(img.numpy(), gt.numpy(), torch.float64)
    ll1, s, loss, g_loss = _host_loss(emu_loss, img.numpy(), gt.numpy(), 0.2)
    for got, want in zip((ll1, s, loss), o[:3]):
        assert abs(got - want) < 5e-6
    assert TL._rel(g_loss, o[3]) < TL.RTOL_GRAD


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 257, 1100])
def test_knn_kernel_on_host_matches_oracle(emu_loss, n):
    from oracle.knn_oracle import dist2_knn3
    gen = torch.Generator().manual_seed(n)
    pts = torch.randn(n, 3, generator=gen)
    if n >= 5:
        pts[3] = pts[1]
    out = torch.full((n,), float("nan"))
    emu_loss.emu_knn3(_p(pts), n, _p(out))
    assert np.allclose(out.numpy(), dist2_knn3(pts.numpy()), rtol=1e-5, atol=1e-7)


# 130 tiles (one partition pass); 5000 Gaussians = two radix tiles per pass; rectangles of more than 64 tiles (no survivor
# mask: the per-row path of the emission); 475 tiles = two partition passes
@pytest.mark.parametrize("P,W,H,smed,yaw", [(1500, 203, 149, 0.04, 8.0), (5000, 96, 64, 0.05, 0.0), (400, 330, 80, 0.3, -4.0),
                                            (2500, 400, 300, 0.03, 3.0)])
def test_binning_on_host_matches_oracle_lists(emu_pre, P, W, H, smed, yaw):
    """scgaussian_b200/csrc/binning.cu on the host, fed by the emulated preprocess: depth sort (4 onesweep passes with
    decoupled look-back), chained scan, instance emission, tile partition + ranges (SURVEY.md section 8a rows a10-a13,
    Appendix A.6 / A.7).  Exact where the algorithm is exact (stable depth order, prefix sums, R), and against the C
    oracle's per-tile lists the same way the GPU stage test does: ours = the reference's lists minus pairs in which no
    pixel can reach alpha >= 1/255, in the same order."""
    from oracle import torch_oracle as O
    from tests import util
    case, t, view, g = _host_scene(P, W, H, 3, seed=P + 2, scale_median=smed, w2c=O.yaw_w2c(yaw), z_shift=-1.9)
    geom, gptr, gv = _geometry(emu_pre, P)
    radii = torch.zeros(P, dtype=torch.int32)
    emu_pre.emu_preprocess_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(radii.data_ptr()))
    touched = gv(1, P * 4, torch.int32).numpy().astype(np.int64).copy()
    emu_pre.emu_depth_sort_and_scan(C.byref(view), C.byref(g), C.c_void_p(gptr))
    keys_sorted = gv(4, P * 4, torch.int32).numpy().view(np.uint32).copy()      # buffer 0 holds the sorted keys after 4 passes
    order = gv(5, P * 4, torch.int32).numpy().astype(np.int64).copy()
    rec = gv(0, P * 48, torch.float32).view(P, 12).numpy()
    zv = (t["means3D"] @ t["viewmatrix"][:3, 2] + t["viewmatrix"][3, 2]).numpy()
    # the keys the sort started from: record depth bits in front of the near plane, CULLED behind it
    emu_lib_keys = np.where(radii.numpy() > 0, rec[:, 6].view(np.uint32), 0)
    assert np.array_equal(np.sort(order), np.arange(P))
    assert (np.diff(keys_sorted.astype(np.int64)) >= 0).all()
    vis = radii.numpy() > 0
    assert np.array_equal(keys_sorted[np.isin(order, np.nonzero(vis)[0])], emu_lib_keys[order][np.isin(order, np.nonzero(vis)[0])])
    ties = np.diff(keys_sorted.astype(np.int64)) == 0
    assert (np.diff(order)[ties] > 0).all()                                       # stable: equal depths keep index order
    emu_pre.emu_offsets_offset.restype = emu_pre.emu_status_offset.restype = C.c_size_t
    base = gptr - geom.data_ptr()
    o_off, s_off = int(emu_pre.emu_offsets_offset(P)), int(emu_pre.emu_status_offset(P))
    offsets = geom[base + o_off: base + o_off + P * 4].view(torch.int32).numpy().astype(np.int64)
    status = geom[base + s_off: base + s_off + 16].view(torch.int64).numpy()
    assert np.array_equal(offsets, np.cumsum(touched[order]))
    R = int(status[0])
    assert R == int(touched.sum()) and R > 0 and int(status[1]) == 0

    emu_pre.emu_binning_bytes.restype = C.c_size_t
    binning = torch.zeros(int(emu_pre.emu_binning_bytes(W, H, C.c_int64(R))) + 64, dtype=torch.uint8)
    bptr = binning.data_ptr() + (-binning.data_ptr()) % 64
    emu_pre.emu_emit_and_partition(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(bptr), C.c_int64(R))
    boff = (C.c_size_t * 4)()
    emu_pre.emu_binning_offsets(W, H, C.c_int64(R), boff)
    bb = bptr - binning.data_ptr()
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    rg = binning[bb + boff[0]: bb + boff[0] + tiles * 8].view(torch.int32).view(tiles, 2).numpy().astype(np.int64).copy()
    rg[rg[:, 1] == 0] = 0                                                          # empty tiles are stored as (0xffffffff, 0)
    pl = binning[bb + boff[1]: bb + boff[1] + R * 4].view(torch.int32).numpy().astype(np.int64)
    assert int(geom[base + s_off: base + s_off + 16].view(torch.int64)[1]) == 0   # no overflow

    co, (c2, r2, d2, a2), _ = util.run_c_oracle(case)
    st = co.state()
    n_rad = int((radii.numpy() != r2).sum())
    ne = rg[:, 1] > rg[:, 0]
    assert rg[ne, 0][0] == 0 and rg[ne, 1][-1] == R and np.array_equal(rg[ne, 0][1:], rg[ne, 1][:-1])
    gx = (W + 15) // 16
    rank = np.empty(P, dtype=np.int64)
    rank[order] = np.arange(P)
    kept = dropped = 0
    for tile in range(tiles):
        mine = pl[rg[tile, 0]:rg[tile, 1]]
        assert (np.diff(rank[mine]) > 0).all(), f"tile {tile}: not in depth order"
        ref = st["point_list"][st["ranges"][tile, 0]:st["ranges"][tile, 1]].astype(np.int64)
        if n_rad == 0:
            pos = {gid: i for i, gid in enumerate(ref)}
            idx = np.array([pos[gid] for gid in mine], dtype=np.int64)           # KeyError = not a subset of the reference's list
            assert (np.diff(idx) > 0).all(), f"tile {tile}: order differs from the reference's"
        gone = np.setdiff1d(ref, mine)
        kept += len(mine)
        dropped += len(gone)
        if len(gone):
            ty, tx = divmod(tile, gx)
            ys, xs = np.meshgrid(np.arange(ty * 16, ty * 16 + 16), np.arange(tx * 16, tx * 16 + 16), indexing="ij")
            dx = st["means2D"][gone, 0][:, None, None] - xs[None]
            dy = st["means2D"][gone, 1][:, None, None] - ys[None]
            cn = st["conic"][gone].astype(np.float64)
            power = -0.5 * (cn[:, 0, None, None] * dx * dx + cn[:, 2, None, None] * dy * dy) - cn[:, 1, None, None] * dx * dy
            al = t["opacities"].numpy()[gone, 0][:, None, None] * np.exp(power)
            assert ((al < 1.0 / 255.0) | (power > 0)).all(), f"tile {tile}: a contributing pair was culled"
    assert kept == R and R <= co.num_rendered and dropped == co.num_rendered - R or n_rad > 0
    rect = gv(2, P * 8, torch.int32).view(P, 2).numpy().astype(np.int64)
    area = ((rect[:, 1] & 0xFFFF) - (rect[:, 0] & 0xFFFF)) * ((rect[:, 1] >> 16) - (rect[:, 0] >> 16))
    if (W, H) == (330, 80):
        assert (area[vis] > 64).sum() >= 20                                        # the no-mask path was exercised


def _host_forward(emu_pre, case, t, view, g, P, W, H):
    """preprocess -> depth sort -> scan -> emission -> partition -> render forward, all on the host."""
    geom, gptr, gv = _geometry(emu_pre, P)
    radii = torch.zeros(P, dtype=torch.int32)
    emu_pre.emu_preprocess_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(radii.data_ptr()))
    emu_pre.emu_depth_sort_and_scan(C.byref(view), C.byref(g), C.c_void_p(gptr))
    emu_pre.emu_status_offset.restype = emu_pre.emu_binning_bytes.restype = emu_pre.emu_image_bytes.restype = C.c_size_t
    s_off = (gptr - geom.data_ptr()) + int(emu_pre.emu_status_offset(P))
    R = int(geom[s_off: s_off + 8].view(torch.int64)[0])
    binning = torch.zeros(int(emu_pre.emu_binning_bytes(W, H, C.c_int64(R))) + 64, dtype=torch.uint8)
    bptr = binning.data_ptr() + (-binning.data_ptr()) % 64
    emu_pre.emu_emit_and_partition(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(bptr), C.c_int64(R))
    image = torch.zeros(int(emu_pre.emu_image_bytes(W, H)) + 64, dtype=torch.uint8)
    iptr = image.data_ptr() + (-image.data_ptr()) % 64
    color, depth, alpha = (torch.full((c, H, W), float("nan")) for c in (3, 1, 1))
    emu_pre.emu_render_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(bptr), C.c_int64(R), C.c_void_p(iptr),
                               _p(color), _p(depth), _p(alpha))
    return dict(geom=geom, gptr=gptr, gv=gv, radii=radii, R=R, binning=binning, bptr=bptr, image=image, iptr=iptr,
                color=color, depth=depth, alpha=alpha)


# work split "halves,quarters" in percent of the tiles: whole tiles (8 slots per warp), halves (4), quarters (2);
# staging of the records: register-prefetched gathers (0) or the TMA path (1; a synchronous copy in the emulation)
@pytest.mark.parametrize("P,W,H,smed,split,tma", [(700, 100, 70, 0.06, "0,0", 0), (700, 100, 70, 0.06, "100,0", 1),
                                                  (400, 64, 48, 0.15, "0,100", 0), (900, 130, 50, 0.05, "30,30", 1),
                                                  (800, 110, 75, 0.06, "20,20", 2)])      # tma 2: cp.async staging (backward)
def test_whole_operator_on_host_matches_oracle(emu_pre, monkeypatch, P, W, H, smed, split, tma):
    """Every kernel of the operator, forward and backward, executed on the host and
    compared with the C oracle the way the GPU parity tests compare the GPU: images within 1e-4, gradients within 1e-3,
    with the documented allowance for isolated discrete flips."""
    from oracle import torch_oracle as O
    from tests import util
    monkeypatch.setenv("SCGR_FWD_SPLIT", split)
    monkeypatch.setenv("SCGR_BWD_SPLIT", split)
    monkeypatch.setenv("SCGR_TMA", str(tma))
    case, t, view, g = _host_scene(P, W, H, 3, seed=P + 4, scale_median=smed, w2c=O.yaw_w2c(5.0), z_shift=-1.5,
                                   bg=(0.1, 0.2, 0.3))
    f = _host_forward(emu_pre, case, t, view, g, P, W, H)
    grads_up = O.synth_upstream_grads(W, H)
    co, (c2, r2, d2, a2), want = util.run_c_oracle(case, grads=grads_up)
    assert not torch.isnan(f["color"]).any() and not torch.isnan(f["depth"]).any() and not torch.isnan(f["alpha"]).any()
    flips = util.flip_sets(co)
    util.assert_image_close("color", f["color"].numpy(), c2, flips)
    util.assert_image_close("depth", f["depth"].numpy(), d2, flips)
    util.assert_image_close("alpha", f["alpha"].numpy(), a2, flips)
    util.assert_radii_match("radii", f["radii"].numpy(), r2, flips)
    assert f["R"] <= co.num_rendered

    gC, gD, gA = [x.contiguous() for x in grads_up]
    emu_pre.emu_render_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.c_void_p(f["bptr"]), C.c_int64(f["R"]),
                                C.c_void_p(f["iptr"]), _p(gC), _p(gD), _p(gA))
    M = int(t["shs"].shape[1])
    outs = {"means3D": torch.full((P, 3), float("nan")), "means2D": torch.full((P, 3), float("nan")),
            "shs": torch.full((P, M, 3), float("nan")), "opacities": torch.full((P, 1), float("nan")),
            "scales": torch.full((P, 3), float("nan")), "rotations": torch.full((P, 4), float("nan"))}
    sg = L.ScgrGrads(outs["means3D"].data_ptr(), outs["means2D"].data_ptr(), outs["shs"].data_ptr(), None,
                     outs["opacities"].data_ptr(), outs["scales"].data_ptr(), outs["rotations"].data_ptr(), None)
    emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.byref(sg))
    for k, v in outs.items():
        assert not torch.isnan(v).any(), k
        util.assert_grad_close(k, v.numpy(), np.asarray(want[k]).reshape(v.shape), flips)
        assert float(v.abs().max()) > 0.0, k


def test_backward_accumulates_over_views_and_emits_densification_stats_on_host(emu_pre, monkeypatch):
    """ScgrGrads.accumulate / .densification_stats (include/scgr.h): the backward of a second view ADDS its parameter
    gradients to the first one's, leaves Gaussians without gradient untouched, keeps dL/dmean2D per view, and the
    statistics columns hold sum_v |dL/dmean2D_v[:, :2]| * visible_v and sum_v visible_v -- what the reference's
    add_densification_stats accumulates over sequential views (scene/gaussian_model.py:932-934)."""
    from oracle import torch_oracle as O
    monkeypatch.setenv("SCGR_FWD_SPLIT", "0,0")
    monkeypatch.setenv("SCGR_BWD_SPLIT", "0,0")
    P, W, H = 900, 120, 90
    M = 16
    per_view, keep = [], []
    acc = {"means3D": torch.full((P, 3), float("nan")), "shs": torch.full((P, M, 3), float("nan")),
           "opacities": torch.full((P, 1), float("nan")), "scales": torch.full((P, 3), float("nan")),
           "rotations": torch.full((P, 4), float("nan")), "stats": torch.full((P, 2), float("nan"))}
    for v, yaw in enumerate((-6.0, 9.0)):
        case, t, view, g = _host_scene(P, W, H, 3, seed=41, scale_median=0.06, w2c=O.yaw_w2c(yaw), z_shift=-1.2)
        f = _host_forward(emu_pre, case, t, view, g, P, W, H)
        gC, gD, gA = [x.contiguous() for x in O.synth_upstream_grads(W, H, seed=3 + v)]
        emu_pre.emu_render_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.c_void_p(f["bptr"]), C.c_int64(f["R"]),
                                    C.c_void_p(f["iptr"]), _p(gC), _p(gD), _p(gA))
        solo = {k: torch.full_like(a, float("nan")) for k, a in acc.items()}
        m2 = torch.full((P, 3), float("nan"))
        sg = L.ScgrGrads(solo["means3D"].data_ptr(), m2.data_ptr(), solo["shs"].data_ptr(), None, solo["opacities"].data_ptr(),
                         solo["scales"].data_ptr(), solo["rotations"].data_ptr(), None, solo["stats"].data_ptr(),
                         f["radii"].data_ptr(), 0)
        emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.byref(sg))
        vis = (f["radii"] > 0).float()
        want_stats = torch.stack([torch.linalg.vector_norm(m2[:, :2], dim=-1) * vis, vis], 1)
        assert torch.allclose(solo["stats"], want_stats, rtol=1e-6, atol=0) and not torch.isnan(solo["stats"]).any()
        assert (vis.sum() > (solo["opacities"][:, 0] != 0).sum()) and vis.sum() < P      # visible-but-gradient-free and culled both occur
        per_view.append((solo, m2.clone()))
        # the same backward into the shared arrays: view 0 overwrites, view 1 accumulates
        m2b = torch.full((P, 3), float("nan"))
        sg = L.ScgrGrads(acc["means3D"].data_ptr(), m2b.data_ptr(), acc["shs"].data_ptr(), None, acc["opacities"].data_ptr(),
                         acc["scales"].data_ptr(), acc["rotations"].data_ptr(), None, acc["stats"].data_ptr(),
                         f["radii"].data_ptr(), int(v > 0))
        emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.byref(sg))
        assert torch.equal(m2b, m2)                                   # per view, never accumulated
        keep.append((f, t, case))
    for k, a in acc.items():
        want = per_view[0][0][k] + per_view[1][0][k]
        assert not torch.isnan(a).any(), k
        assert torch.allclose(a, want, rtol=1e-6, atol=1e-12), k
    assert float(acc["stats"][:, 1].max()) == 2.0


@pytest.mark.parametrize("split", ["0,0", "40,30"])
def test_packed_and_scalar_forward_blends_are_bit_identical_on_host(emu_pre, monkeypatch, split):
    """render.cu's two forward blends -- scalar, and packed fp32 over the two slot columns of a tile row (FFMA2 / FMUL2
    on sm_100a) -- perform the same per-pixel operations with the same roundings: identical images, n_contrib and
    final_T, bit for bit (whole tiles, halves and quarters)."""
    from oracle import torch_oracle as O
    monkeypatch.setenv("SCGR_FWD_SPLIT", split)
    P, W, H = 1200, 150, 100
    case, t, view, g = _host_scene(P, W, H, 3, seed=52, scale_median=0.07, w2c=O.yaw_w2c(-4.0), z_shift=-1.0, bg=(0.2, 0.1, 0.4))
    outs = []
    for packed in ("0", "1"):
        monkeypatch.setenv("SCGR_FWD_PACKED", packed)
        f = _host_forward(emu_pre, case, t, view, g, P, W, H)
        img = f["image"][f["iptr"] - f["image"].data_ptr():][: 2 * W * H * 4].clone()      # n_contrib + final_T
        outs.append((f["color"].clone(), f["depth"].clone(), f["alpha"].clone(), img))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert float(outs[0][2].max()) > 0.5


def test_overflow_flag_on_host_describes_the_last_emission(emu_pre):
    """The recovery path of SCGR_NEED_CAPACITY at kernel level: an emission into a binning buffer that is too small
    writes nothing and raises the device flag; the re-run with a large enough buffer (stage 1 kept) clears it and
    produces the same lists as a first-time run."""
    from oracle import torch_oracle as O
    P, W, H = 1200, 160, 120
    case, t, view, g = _host_scene(P, W, H, 3, seed=9, scale_median=0.05)
    geom, gptr, gv = _geometry(emu_pre, P)
    radii = torch.zeros(P, dtype=torch.int32)
    emu_pre.emu_preprocess_forward(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(radii.data_ptr()))
    emu_pre.emu_depth_sort_and_scan(C.byref(view), C.byref(g), C.c_void_p(gptr))
    emu_pre.emu_status_offset.restype = emu_pre.emu_binning_bytes.restype = C.c_size_t
    s_off = (gptr - geom.data_ptr()) + int(emu_pre.emu_status_offset(P))
    status = geom[s_off: s_off + 16].view(torch.int64)
    R = int(status[0])
    assert R > 2000 and int(status[1]) == 0

    def run(capacity):
        binning = torch.full((int(emu_pre.emu_binning_bytes(W, H, C.c_int64(capacity))) + 64,), 0xAB, dtype=torch.uint8)
        bptr = binning.data_ptr() + (-binning.data_ptr()) % 64
        emu_pre.emu_emit_and_partition(C.byref(view), C.byref(g), C.c_void_p(gptr), C.c_void_p(bptr), C.c_int64(capacity))
        boff = (C.c_size_t * 4)()
        emu_pre.emu_binning_offsets(W, H, C.c_int64(capacity), boff)
        bb = bptr - binning.data_ptr()
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        n = min(capacity, R)
        return (binning[bb + boff[0]: bb + boff[0] + tiles * 8].view(torch.int32).clone(),
                binning[bb + boff[1]: bb + boff[1] + n * 4].view(torch.int32).clone(),
                binning[bb + boff[2]: bb + boff[2] + n * 4].clone())
    _, _, raw_small = run(R // 2)
    assert int(status[0]) == R and int(status[1]) == 1
    assert (raw_small == 0xAB).all()                       # nothing was written into the undersized buffer
    ranges_a, list_a, _ = run(R)
    assert int(status[1]) == 0                             # the flag describes THIS emission
    ranges_b, list_b, _ = run(R + 777)                     # over-allocated, as the fused protocol does
    assert int(status[1]) == 0
    assert torch.equal(ranges_a, ranges_b) and torch.equal(list_a, list_b)


@pytest.mark.parametrize("mode", ["precomp", "sh_deg1_of_2"])
def test_operator_input_variants_on_host(emu_pre, monkeypatch, mode):
    """The other input paths of the operator (reference gaussian_renderer/__init__.py:61-87) through the host-emulated
    kernels: precomputed colour + covariance (no SH, no scale / rotation), and an SH layout other than 16
    coefficients with the active degree below the maximum (the non-staged SH path)."""
    from oracle import torch_oracle as O
    from tests import util
    monkeypatch.setenv("SCGR_FWD_SPLIT", "50,20")
    monkeypatch.setenv("SCGR_BWD_SPLIT", "50,20")
    monkeypatch.setenv("SCGR_TMA", "1")
    P, W, H = 600, 96, 64
    if mode == "precomp":
        case, t, view, g = _host_scene(P, W, H, 2, seed=31, scale_median=0.08, bg=(0.1, 0.1, 0.4), scale_modifier=1.2)
        gen = torch.Generator().manual_seed(3)
        col = torch.rand(P, 3, generator=gen)
        c3 = O.cov3d_from_scale_rot(case["scales"], case["rotations"], case["scale_modifier"])
        c6 = torch.stack([c3[:, 0, 0], c3[:, 0, 1], c3[:, 0, 2], c3[:, 1, 1], c3[:, 1, 2], c3[:, 2, 2]], -1).contiguous()
        g = L.ScgrGaussians(P, 0, t["means3D"].data_ptr(), t["opacities"].data_ptr(), None, col.data_ptr(), None, None,
                            c6.data_ptr())
        kw = dict(colors_precomp=col, cov3D_precomp=c6)
        keys = ("means3D", "means2D", "opacities", "colors_precomp", "cov3D_precomp")
    else:
        case, t, view, g = _host_scene(P, W, H, 1, seed=32, max_sh_degree=2, scale_median=0.08)      # M = 9, degree 1 active
        kw = {}
        keys = ("means3D", "means2D", "opacities", "shs", "scales", "rotations")
    f = _host_forward(emu_pre, case, t, view, g, P, W, H)
    grads_up = O.synth_upstream_grads(W, H)
    co, (c2, r2, d2, a2), want = util.run_c_oracle(case, grads=grads_up, **kw)
    flips = util.flip_sets(co)
    util.assert_image_close("color", f["color"].numpy(), c2, flips)
    util.assert_image_close("depth", f["depth"].numpy(), d2, flips)
    util.assert_image_close("alpha", f["alpha"].numpy(), a2, flips)
    gC, gD, gA = [x.contiguous() for x in grads_up]
    emu_pre.emu_render_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.c_void_p(f["bptr"]), C.c_int64(f["R"]),
                                C.c_void_p(f["iptr"]), _p(gC), _p(gD), _p(gA))
    M = int(t["shs"].shape[1])
    shapes = {"means3D": (P, 3), "means2D": (P, 3), "shs": (P, M, 3), "colors_precomp": (P, 3), "opacities": (P, 1),
              "scales": (P, 3), "rotations": (P, 4), "cov3D_precomp": (P, 6)}
    outs = {k: torch.full(shapes[k], float("nan")) for k in keys}
    ptr = lambda k: outs[k].data_ptr() if k in outs else None      # noqa: E731
    sg = L.ScgrGrads(ptr("means3D"), ptr("means2D"), ptr("shs"), ptr("colors_precomp"), ptr("opacities"), ptr("scales"),
                     ptr("rotations"), ptr("cov3D_precomp"))
    emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.byref(sg))
    for k, v in outs.items():
        assert not torch.isnan(v).any(), k
        util.assert_grad_close(k, v.numpy(), np.asarray(want[k]).reshape(v.shape), flips)
    if mode != "precomp":
        assert float(outs["shs"][:, 4:].abs().max()) == 0.0       # rows beyond the active degree stay zero


def test_whole_operator_on_host_random_scenes(emu_pre, monkeypatch):
    """A seeded sweep over odd shapes -- one or two Gaussians, single-tile and ragged images, huge and tiny splats,
    every SH degree, scenes where nothing is rendered -- forward and backward against the C oracle.  (100 such
    configurations were also run under AddressSanitizer while this was written: no failure, no out-of-bounds access.)"""
    import random
    from oracle import torch_oracle as O
    from tests import util
    rng = random.Random(5)
    seen_empty = False
    for it in range(14):
        P = rng.choice([1, 2, 33, 200, 700]); W = rng.choice([16, 17, 64, 100, 203]); H = rng.choice([16, 31, 48, 70])      # noqa: E702
        deg = rng.choice([0, 1, 2, 3]); smed = rng.choice([0.01, 0.05, 0.15, 0.6]); yaw = rng.choice([0.0, 8.0, -25.0])      # noqa: E702
        split = rng.choice(["0,0", "100,0", "0,100", "40,40"]); tma = rng.choice([0, 1]); zs = rng.choice([0.0, -1.9, -5.0])  # noqa: E702
        mod = rng.choice([1.0, 0.7, 2.0]); bg = rng.choice([(0.0, 0.0, 0.0), (0.1, 0.2, 0.3)])                              # noqa: E702
        monkeypatch.setenv("SCGR_FWD_SPLIT", split)
        monkeypatch.setenv("SCGR_BWD_SPLIT", split)
        monkeypatch.setenv("SCGR_TMA", str(tma))
        cfg = dict(P=P, W=W, H=H, deg=deg, smed=smed, yaw=yaw, split=split, tma=tma, zs=zs, mod=mod, bg=bg)
        case, t, view, g = _host_scene(P, W, H, deg, seed=it + 100, scale_median=smed, w2c=O.yaw_w2c(yaw), z_shift=zs, bg=bg,
                                       scale_modifier=mod)
        f = _host_forward(emu_pre, case, t, view, g, P, W, H)
        gu = O.synth_upstream_grads(W, H)
        co, (c2, r2, d2, a2), want = util.run_c_oracle(case, grads=gu)
        seen_empty |= f["R"] == 0
        flips = util.flip_sets(co)
        for name, got, ref in (("color", f["color"], c2), ("depth", f["depth"], d2), ("alpha", f["alpha"], a2)):
            util.assert_image_close(f"{name} {cfg}", got.numpy(), ref, flips)
        util.assert_radii_match(f"radii {cfg}", f["radii"].numpy(), r2, flips)
        gC, gD, gA = [x.contiguous() for x in gu]
        emu_pre.emu_render_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.c_void_p(f["bptr"]), C.c_int64(f["R"]),
                                    C.c_void_p(f["iptr"]), _p(gC), _p(gD), _p(gA))
        M = int(t["shs"].shape[1])
        outs = {"means3D": torch.full((P, 3), float("nan")), "means2D": torch.full((P, 3), float("nan")),
                "shs": torch.full((P, M, 3), float("nan")), "opacities": torch.full((P, 1), float("nan")),
                "scales": torch.full((P, 3), float("nan")), "rotations": torch.full((P, 4), float("nan"))}
        sg = L.ScgrGrads(outs["means3D"].data_ptr(), outs["means2D"].data_ptr(), outs["shs"].data_ptr(), None,
                         outs["opacities"].data_ptr(), outs["scales"].data_ptr(), outs["rotations"].data_ptr(), None)
        emu_pre.emu_preprocess_backward(C.byref(view), C.byref(g), C.c_void_p(f["gptr"]), C.byref(sg))
        for k, v in outs.items():
            assert not torch.isnan(v).any(), (k, cfg)
            util.assert_grad_close(f"{k} {cfg}", v.numpy(), np.asarray(want[k]).reshape(v.shape), flips)
    assert seen_empty                                  # the sweep includes a view in which nothing is rendered


# ------------------------------------------------------------------------------------------------------------------
# The whole library behind its real C entry points (capi.cu included), on the host.  Every exported symbol of this
# build is renamed scgr_* -> emu_scgr_*: only a test can bind it.
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_abi():
    from tests.emulation import build
    try:
        path = build.build_full()
    except Exception as e:      # pragma: no cover
        pytest.skip(f"host emulation library not buildable here: {e}")
    lib = C.CDLL(path)
    ns = type("EmuAbi", (), {})()
    for name, (res, args) in L.SYMBOLS.items():
        fn = getattr(lib, "emu_" + name)          # AttributeError: an entry point of include/scgr.h is missing
        fn.restype, fn.argtypes = res, args
        setattr(ns, name, fn)
    ns.path = path
    return ns


def _aligned(nbytes, fill=0):
    buf = torch.full((int(nbytes) + 256,), fill, dtype=torch.uint8)
    return buf, buf.data_ptr() + (-buf.data_ptr()) % 256


def test_c_abi_on_host_protocols_and_overflow_recovery(emu_abi, monkeypatch):
    """scgr_forward (both stages, host wait for R on the status word), SCGR_NEED_CAPACITY when the pre-sized binning
    buffer is too small, completion by scgr_forward_render with stage 1 kept, scgr_backward -- the C-level orchestration
    of scgaussian_b200/csrc/capi.cu, against the C oracle.  After the recovery the device status reads {R, 0}."""
    import subprocess
    from oracle import torch_oracle as O
    from tests import util
    exported = subprocess.run(["nm", "-D", "--defined-only", emu_abi.path], capture_output=True, text=True).stdout
    assert " emu_scgr_forward" in exported and " scgr_forward" not in exported       # cannot stand in for libscgr.so
    monkeypatch.setenv("SCGR_FWD_SPLIT", "0,0")
    monkeypatch.setenv("SCGR_BWD_SPLIT", "0,0")
    P, W, H = 900, 120, 90
    case, t, view, g = _host_scene(P, W, H, 3, seed=77, scale_median=0.06, bg=(0.1, 0.2, 0.3), z_shift=-1.5)
    grads_up = O.synth_upstream_grads(W, H)
    co, (c2, r2, d2, a2), want = util.run_c_oracle(case, grads=grads_up)

    def forward(capacity):
        geom, gptr = _aligned(emu_abi.scgr_geometry_bytes(P))
        image, iptr = _aligned(emu_abi.scgr_image_bytes(W, H))
        binning, bptr = _aligned(emu_abi.scgr_binning_bytes(P, W, H, capacity))
        radii = torch.zeros(P, dtype=torch.int32)
        color, depth, alpha = (torch.full((c, H, W), float("nan")) for c in (3, 1, 1))
        status = torch.zeros(2, dtype=torch.int64)
        rc = emu_abi.scgr_forward(C.byref(view), C.byref(g), gptr, radii.data_ptr(), bptr, capacity, iptr, color.data_ptr(),
                                  depth.data_ptr(), alpha.data_ptr(), status.data_ptr(), None)
        return rc, dict(geom=geom, gptr=gptr, image=image, iptr=iptr, binning=binning, bptr=bptr, radii=radii, color=color,
                        depth=depth, alpha=alpha, status=status, capacity=capacity)

    rc, big = forward(co.num_rendered + 5000)                       # ample capacity: one call
    assert rc == 0, emu_abi.scgr_last_error()
    R = int(big["status"][0])
    assert 0 < R <= co.num_rendered
    flips = util.flip_sets(co)
    util.assert_image_close("color", big["color"].numpy(), c2, flips)
    util.assert_image_close("depth", big["depth"].numpy(), d2, flips)
    util.assert_image_close("alpha", big["alpha"].numpy(), a2, flips)

    rc, small = forward(R // 3)                                     # too small: stage 1 done, stage 2 refused
    assert rc == L.NEED_CAPACITY and int(small["status"][0]) == R
    assert torch.isnan(small["color"]).all()                        # nothing was rendered
    dv = L.ScgrDebugViews()
    assert emu_abi.scgr_debug_views(P, W, H, R // 3, small["gptr"], small["bptr"], small["iptr"], C.byref(dv)) == 0
    dev_status = (C.c_int64 * 2).from_address(C.cast(dv.num_rendered, C.c_void_p).value)
    assert (dev_status[0], dev_status[1]) == (R, 1)                 # the refused emission raised the flag
    binning, bptr = _aligned(emu_abi.scgr_binning_bytes(P, W, H, R))
    rc = emu_abi.scgr_forward_render(C.byref(view), C.byref(g), small["gptr"], bptr, R, small["iptr"], small["color"].data_ptr(),
                                     small["depth"].data_ptr(), small["alpha"].data_ptr(), small["status"].data_ptr(), None)
    assert rc == 0, emu_abi.scgr_last_error()
    assert (dev_status[0], dev_status[1]) == (R, 0)                 # ... and the completed one cleared it
    assert int(small["status"][0]) == R and int(small["status"][1]) == 0
    for k in ("color", "depth", "alpha"):
        assert torch.equal(small[k], big[k]), k                     # bit-identical to the run that never overflowed
    assert torch.equal(small["radii"], big["radii"])

    gC, gD, gA = [x.contiguous() for x in grads_up]
    M = int(t["shs"].shape[1])
    outs = {"means3D": torch.full((P, 3), float("nan")), "means2D": torch.full((P, 3), float("nan")),
            "shs": torch.full((P, M, 3), float("nan")), "opacities": torch.full((P, 1), float("nan")),
            "scales": torch.full((P, 3), float("nan")), "rotations": torch.full((P, 4), float("nan"))}
    sg = L.ScgrGrads(outs["means3D"].data_ptr(), outs["means2D"].data_ptr(), outs["shs"].data_ptr(), None,
                     outs["opacities"].data_ptr(), outs["scales"].data_ptr(), outs["rotations"].data_ptr(), None)
    rc = emu_abi.scgr_backward(C.byref(view), C.byref(g), small["gptr"], bptr, R, small["iptr"], gC.data_ptr(), gD.data_ptr(),
                               gA.data_ptr(), C.byref(sg), None)
    assert rc == 0, emu_abi.scgr_last_error()
    for k, v in outs.items():
        util.assert_grad_close(k, v.numpy(), np.asarray(want[k]).reshape(v.shape), flips)

    # P = 0 through the fused entry point: zero images, not background-filled
    g0 = L.ScgrGaussians(0, 16, None, None, None, None, None, None, None)
    color = torch.full((3, H, W), float("nan"))
    depth, alpha = torch.full((1, H, W), float("nan")), torch.full((1, H, W), float("nan"))
    status = torch.full((2,), -1, dtype=torch.int64)
    image, iptr = _aligned(emu_abi.scgr_image_bytes(W, H))
    geom, gptr = _aligned(emu_abi.scgr_geometry_bytes(0))
    binning0, bptr0 = _aligned(emu_abi.scgr_binning_bytes(0, W, H, 16))
    rc = emu_abi.scgr_forward(C.byref(view), C.byref(g0), gptr, None, bptr0, 16, iptr, color.data_ptr(), depth.data_ptr(),
                              alpha.data_ptr(), status.data_ptr(), None)
    assert rc == 0, emu_abi.scgr_last_error()
    assert int(status[0]) == 0 and float(color.abs().max()) == 0.0 and float(alpha.abs().max()) == 0.0

    # argument errors come back as a status + message, never as an exception across the boundary
    bad = L.ScgrGaussians(P, 16, t["means3D"].data_ptr(), t["opacities"].data_ptr(), None, None, t["scales"].data_ptr(),
                          t["rotations"].data_ptr(), None)              # neither shs nor colors_precomp
    rc = emu_abi.scgr_forward(C.byref(view), C.byref(bad), big["gptr"], big["radii"].data_ptr(), big["bptr"], 10, big["iptr"],
                              color.data_ptr(), depth.data_ptr(), alpha.data_ptr(), status.data_ptr(), None)
    assert rc == 1 and b"exactly one of shs / colors_precomp" in emu_abi.scgr_last_error()


@pytest.mark.parametrize("n0_of", ["mid", "zero", "all", "one"])
def test_split_sh_layout_on_host_is_bit_identical_to_the_assembled_one(emu_abi, monkeypatch, n0_of):
    """SURVEY 8f row f2, second half (include/scgr.h: ScgrGaussians.sh_dc / sh_rest): the operator reading the SH rows
    straight from the hybrid model's four arrays (reference scene/gaussian_model.py:131-140 cats them on every render)
    gives the same bits as the assembled [P,16,3] input -- images, radii, and every gradient, the dL/dfeatures_* rows being
    the slices of dL/dshs; accumulate mode included."""
    from oracle import torch_oracle as O
    monkeypatch.setenv("SCGR_FWD_SPLIT", "0,0")
    monkeypatch.setenv("SCGR_BWD_SPLIT", "0,0")
    P, W, H = 700, 120, 90
    case, t, view, g = _host_scene(P, W, H, 3, seed=5, scale_median=0.06, bg=(0.1, 0.2, 0.3), z_shift=-1.7)
    n0 = {"mid": 333, "zero": 0, "all": P, "one": 1}[n0_of]          # 333: a set boundary inside a 128-Gaussian block
    shs = t["shs"]
    # each array in its own allocation, sized exactly: rows of 3 / 45 floats, no padding
    dc = [shs[:n0, :1].clone().contiguous(), shs[n0:, :1].clone().contiguous()]
    rest = [shs[:n0, 1:].clone().contiguous(), shs[n0:, 1:].clone().contiguous()]
    gs = L.ScgrGaussians(P, 16, t["means3D"].data_ptr(), t["opacities"].data_ptr(), None, None, t["scales"].data_ptr(),
                         t["rotations"].data_ptr(), None)
    for k in range(2):
        if dc[k].shape[0]:
            gs.sh_dc[k], gs.sh_rest[k] = dc[k].data_ptr(), rest[k].data_ptr()
    gs.sh_n0 = n0
    gC, gD, gA = [x.contiguous() for x in O.synth_upstream_grads(W, H)]

    def run(gg, split, accumulate):
        cap = 60000
        geom, gptr = _aligned(emu_abi.scgr_geometry_bytes(P))
        image, iptr = _aligned(emu_abi.scgr_image_bytes(W, H))
        binning, bptr = _aligned(emu_abi.scgr_binning_bytes(P, W, H, cap))
        radii = torch.zeros(P, dtype=torch.int32)
        color, depth, alpha = (torch.full((c, H, W), float("nan")) for c in (3, 1, 1))
        status = torch.zeros(2, dtype=torch.int64)
        rc = emu_abi.scgr_forward(C.byref(view), C.byref(gg), gptr, radii.data_ptr(), bptr, cap, iptr, color.data_ptr(),
                                  depth.data_ptr(), alpha.data_ptr(), status.data_ptr(), None)
        assert rc == 0, emu_abi.scgr_last_error()
        fill = 0.25 if accumulate else float("nan")
        outs = {"means3D": torch.full((P, 3), fill), "means2D": torch.full((P, 3), float("nan")),
                "opacities": torch.full((P, 1), fill), "scales": torch.full((P, 3), fill),
                "rotations": torch.full((P, 4), fill)}
        sg = L.ScgrGrads(outs["means3D"].data_ptr(), outs["means2D"].data_ptr(), None, None, outs["opacities"].data_ptr(),
                         outs["scales"].data_ptr(), outs["rotations"].data_ptr(), None)
        sg.accumulate = int(accumulate)
        if split:
            outs["dc"] = [torch.full_like(x, fill) for x in dc]
            outs["rest"] = [torch.full_like(x, fill) for x in rest]
            for k in range(2):
                if dc[k].shape[0]:
                    sg.dL_dsh_dc[k], sg.dL_dsh_rest[k] = outs["dc"][k].data_ptr(), outs["rest"][k].data_ptr()
        else:
            outs["shs"] = torch.full((P, 16, 3), fill)
            sg.dL_dshs = outs["shs"].data_ptr()
        rc = emu_abi.scgr_backward(C.byref(view), C.byref(gg), gptr, bptr, cap, iptr, gC.data_ptr(), gD.data_ptr(),
                                   gA.data_ptr(), C.byref(sg), None)
        assert rc == 0, emu_abi.scgr_last_error()
        return dict(color=color, depth=depth, alpha=alpha, radii=radii, R=int(status[0])), outs

    for accumulate in (False, True):
        img_a, out_a = run(g, False, accumulate)
        img_s, out_s = run(gs, True, accumulate)
        assert img_a["R"] == img_s["R"] > 0
        for k in ("color", "depth", "alpha", "radii"):
            assert torch.equal(img_a[k], img_s[k]), k
        # (render_backward sums the per-Gaussian screen-space gradients with atomics in arrival order: two runs of the SAME
        # layout differ in the last bits too, so the gradients are compared to rounding, the forward bit for bit)
        def same(name, a, b):
            assert a.shape == b.shape and not torch.isnan(a).any() and not torch.isnan(b).any(), name
            if a.numel():
                assert float((a - b).abs().max()) <= 2e-5 * float(a.abs().max()) + 1e-12, name
                assert torch.equal(a == 0, b == 0), name          # rows without gradient are exact zeros in both layouts

        for k in ("means3D", "means2D", "opacities", "scales", "rotations"):
            same(k, out_a[k], out_s[k])
        sh_ref = out_a["shs"] - (0.25 if accumulate else 0.0)
        assert float(sh_ref.abs().max()) > 0
        same("dc0", out_a["shs"][:n0, :1], out_s["dc"][0]); same("dc1", out_a["shs"][n0:, :1], out_s["dc"][1])
        same("rest0", out_a["shs"][:n0, 1:], out_s["rest"][0]); same("rest1", out_a["shs"][n0:, 1:], out_s["rest"][1])

    # argument errors of the split layout
    bad = L.ScgrGaussians(P, 9, t["means3D"].data_ptr(), t["opacities"].data_ptr(), None, None, t["scales"].data_ptr(),
                          t["rotations"].data_ptr(), None)
    bad.sh_dc[1], bad.sh_rest[1], bad.sh_n0 = dc[1].data_ptr() or dc[0].data_ptr(), rest[1].data_ptr() or rest[0].data_ptr(), 0
    status = torch.zeros(2, dtype=torch.int64)
    rc = emu_abi.scgr_forward(C.byref(view), C.byref(bad), 256, 256, 256, 10, 256, 256, 256, 256, status.data_ptr(), None)
    assert rc == 1 and b"need 16 coefficients" in emu_abi.scgr_last_error()
