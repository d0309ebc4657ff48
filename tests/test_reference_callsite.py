"""The reference's own call site against the drop-in package.  Needs the reference's Python files: /root/reference in
the build container, or the staged copy under oracle/_ref/reference that __graft_entry__.build() makes there (git-
ignored, never committed, travels to the GPU box with the snapshot) -- so the `-m gpu` tests at the bottom run the
reference's UNMODIFIED gaussian_renderer.render() and GaussianModel accessors on a B200 against libscgr.so.

`gaussian_renderer/__init__.py` is imported UNMODIFIED with this repo on sys.path, so its line 15
(`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`) resolves to
the B200 operator, and its `render()` is run on a small hybrid-style model up to the operator call:
the 12 settings fields (:38-51), the constructor (:53), the keyword arguments and the 4-tuple unpacking of
the call (:100-108) and the returned dict (:113-118) are exercised with the reference's own code.
No GPU here, so the operator's compute is replaced by a recorder for the duration of the call (the real
operator refuses CPU tensors: there is no CPU path) -- what is checked is the boundary, not the pixels.
The reference's unrelated, uninstalled dependencies (plyfile, simple_knn, pytorch3d, skimage, imageio,
matplotlib) are stubbed for the import."""
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((d for d in (os.environ.get("SCGR_REFERENCE_DIR", ""), "/root/reference",
                        os.path.join(ROOT, "oracle", "_ref", "reference"))
            if d and os.path.isdir(os.path.join(d, "gaussian_renderer"))), "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gaussian_renderer")),
                                reason="the reference's Python files are neither at /root/reference nor staged under oracle/_ref")

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


@pytest.fixture()
def reference_renderer(monkeypatch):
    finder = _StubFinder()
    sys.meta_path.append(finder)         # last: only packages that are really missing get stubbed
    monkeypatch.syspath_prepend(REF)
    monkeypatch.syspath_prepend(ROOT)
    before = set(sys.modules)
    try:
        import gaussian_renderer
        yield gaussian_renderer
    finally:
        sys.meta_path.remove(finder)
        for name in set(sys.modules) - before:     # forget the stubs and the reference's own modules, nothing else
            if isinstance(sys.modules[name], _Stub) or \
                    name.split(".")[0] in ("gaussian_renderer", "scene", "utils", "arguments", "lpipsPyTorch"):
                sys.modules.pop(name, None)


def test_reference_render_reaches_the_drop_in_operator(reference_renderer, monkeypatch):
    import diff_gaussian_rasterization as D
    from scgaussian_b200 import rasterizer as R
    G = reference_renderer
    assert G.GaussianRasterizer is R.GaussianRasterizer is D.GaussianRasterizer          # reference line 15
    assert G.GaussianRasterizationSettings is R.GaussianRasterizationSettings

    P, H, W = 7, 12, 20
    g = torch.Generator().manual_seed(0)

    class Cam:                                   # what reference scene/cameras.py:54-63 exposes
        FoVx, FoVy = 1.0, 0.7
        image_height, image_width = H, W
        world_view_transform = torch.eye(4)
        full_proj_transform = torch.eye(4)
        camera_center = torch.zeros(3)

    class PC:                                    # accessors of reference scene/gaussian_model.py:105-152
        active_sh_degree = 2
        max_sh_degree = 3
        get_xyz = torch.randn(P, 3, generator=g).requires_grad_(True)
        get_opacity = torch.rand(P, 1, generator=g)
        get_scaling = torch.rand(P, 3, generator=g)
        get_rotation = torch.nn.functional.normalize(torch.randn(P, 4, generator=g))
        get_features = torch.randn(P, 16, 3, generator=g)

    class Pipe:                                  # reference arguments/__init__.py:66-68
        convert_SHs_python = False
        compute_cov3D_python = False
        debug = False

    # reference line 28 hard-codes device="cuda"; on this CPU-only box map it to the tensors' own device
    real_zeros_like = torch.zeros_like
    monkeypatch.setattr(torch, "zeros_like", lambda t, **k: real_zeros_like(t, **{**k, "device": t.device}))

    seen = {}

    def recorder(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                 cov3D_precomp=None):
        seen.update(settings=self.raster_settings, means3D=means3D, means2D=means2D, opacities=opacities, shs=shs,
                    colors_precomp=colors_precomp, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
        s = self.raster_settings
        color = means3D.sum() * torch.ones(3, s.image_height, s.image_width)           # keeps autograd connected
        return color, torch.arange(P, dtype=torch.int32) % 3, torch.ones(1, s.image_height, s.image_width), \
            torch.zeros(1, s.image_height, s.image_width)

    monkeypatch.setattr(R.GaussianRasterizer, "forward", recorder)
    bg = torch.tensor([0.1, 0.2, 0.3])
    out = G.render(Cam(), PC(), Pipe(), bg, scaling_modifier=0.5)

    s = seen["settings"]                          # reference lines 38-51
    assert isinstance(s, R.GaussianRasterizationSettings)
    assert (s.image_height, s.image_width, s.sh_degree, s.prefiltered, s.debug) == (H, W, 2, False, False)
    assert s.scale_modifier == 0.5 and s.bg is bg
    assert abs(s.tanfovx - math.tan(0.5)) < 1e-12 and abs(s.tanfovy - math.tan(0.35)) < 1e-12
    assert s.viewmatrix is Cam.world_view_transform and s.projmatrix is Cam.full_proj_transform
    assert s.campos is Cam.camera_center
    # reference lines 100-108: SH + scale/rotation path, nothing precomputed
    assert seen["shs"] is PC.get_features and seen["colors_precomp"] is None
    assert seen["scales"] is PC.get_scaling and seen["rotations"] is PC.get_rotation and seen["cov3D_precomp"] is None
    assert seen["means3D"] is PC.get_xyz and seen["opacities"] is PC.get_opacity
    assert seen["means2D"].shape == (P, 3) and seen["means2D"].requires_grad               # the gradient hook (:28-32)
    # reference lines 113-118: the dict train.py / render.py consume
    assert set(out) >= {"render", "rendered_depth", "rendered_alpha", "viewspace_points", "visibility_filter", "radii"}
    assert out["render"].shape == (3, H, W) and out["rendered_depth"].shape == (1, H, W)
    assert out["viewspace_points"] is seen["means2D"]
    assert torch.equal(out["visibility_filter"], out["radii"] > 0)


def test_the_real_operator_refuses_the_cpu_call(reference_renderer, monkeypatch):
    """Same call without the recorder: the product has no CPU path and says so."""
    from scgaussian_b200 import ScgrError
    G = reference_renderer
    real_zeros_like = torch.zeros_like
    monkeypatch.setattr(torch, "zeros_like", lambda t, **k: real_zeros_like(t, **{**k, "device": t.device}))

    class Cam:
        FoVx, FoVy = 1.0, 0.7
        image_height, image_width = 8, 8
        world_view_transform = torch.eye(4)
        full_proj_transform = torch.eye(4)
        camera_center = torch.zeros(3)

    class PC:
        active_sh_degree = 0
        max_sh_degree = 0
        get_xyz = torch.zeros(3, 3)
        get_opacity = torch.ones(3, 1)
        get_scaling = torch.ones(3, 3)
        get_rotation = torch.tensor([[1.0, 0, 0, 0]]).repeat(3, 1)
        get_features = torch.zeros(3, 1, 3)

    class Pipe:
        convert_SHs_python = compute_cov3D_python = debug = False

    with pytest.raises(ScgrError, match="CUDA"):
        G.render(Cam(), PC(), Pipe(), torch.zeros(3))


def test_reference_optimizer_surgery_runs_on_the_fused_adam(reference_renderer):
    """The reference's densification edits its optimizers through `param_groups` / `state` (reference
    scene/gaussian_model.py:758-843: replace_tensor_to_optimizer, _prune_optimizer, cat_tensors_to_optimizer).
    Those methods, UNMODIFIED, are run here on scgaussian_b200.optim.Adam and on
    torch.optim.Adam side by side: same parameters, same state afterwards.  (Host logic only: no step() on the CPU.)"""
    from scene.gaussian_model import GaussianModel      # importable once the fixture has installed the stubs
    from scgaussian_b200 import optim
    g = torch.Generator().manual_seed(4)

    def build(cls):
        ps = {"zval": torch.nn.Parameter(torch.randn(6, 1, generator=torch.Generator().manual_seed(1))),
              "opacity": torch.nn.Parameter(torch.randn(6, 1, generator=torch.Generator().manual_seed(2)))}
        opt = cls([{"params": [p], "lr": 0.01, "name": n} for n, p in ps.items()], lr=0.0, eps=1e-15)
        for n, p in ps.items():       # a state as after one step (torch's key set)
            opt.state[p] = {"step": torch.tensor(1.0), "exp_avg": torch.full_like(p, 0.5),
                            "exp_avg_sq": torch.full_like(p, 0.25)}
        return opt

    pc = GaussianModel(3)
    mask = torch.tensor([True, False, True, True, False, True])
    ext = {"zval": torch.randn(2, 1, generator=g), "opacity": torch.randn(2, 1, generator=g)}
    results = []
    for cls in (optim.Adam, torch.optim.Adam):
        opt = build(cls)
        a = pc._prune_optimizer(mask, opt)
        assert a["zval"].shape == (4, 1) and opt.state[a["zval"]]["exp_avg"].shape == (4, 1)
        b = pc.cat_tensors_to_optimizer({k: v.clone() for k, v in ext.items()}, opt)
        assert b["opacity"].shape == (6, 1) and float(opt.state[b["opacity"]]["exp_avg"][-1]) == 0.0
        c = pc.replace_tensor_to_optimizer(torch.full((6, 1), -2.0), "opacity", opt)
        st = opt.state[c["opacity"]]
        assert float(st["exp_avg"].abs().max()) == 0.0 and float(st["step"]) == 1.0
        assert [grp["name"] for grp in opt.param_groups] == ["zval", "opacity"]
        results.append({grp["name"]: (grp["params"][0].detach().clone(), opt.state[grp["params"][0]]["exp_avg"].clone())
                        for grp in opt.param_groups})
    for name in ("zval", "opacity"):
        assert torch.equal(results[0][name][0], results[1][name][0])
        assert torch.equal(results[0][name][1], results[1][name][1])


class _NumpyDouble:
    """Test double of the two data-movement entry points of libscgr.so (scgr_gather_rows, scgr_copy_segments) on HOST
    pointers, honouring the same ctypes tables.  The build container has no GPU and the product has no CPU path: this
    double exists so that the HOST LOGIC of scgaussian_b200/densify.py can be run next to the reference's own methods;
    the kernels themselves are compared with torch on the GPU (tests/test_model.py)."""
    calls = 0

    def scgr_gather_rows(self, table, n_arrays, index_ptr, n_out, stream):
        import ctypes as C
        import numpy as np
        _NumpyDouble.calls += 1
        idx = np.ctypeslib.as_array((C.c_int64 * n_out).from_address(index_ptr))
        for a in range(n_arrays):
            row = table[a].row_floats
            src = np.ctypeslib.as_array((C.c_float * ((int(idx.max()) + 1) * row)).from_address(table[a].src)).reshape(-1, row)
            dst = np.ctypeslib.as_array((C.c_float * (n_out * row)).from_address(table[a].dst)).reshape(-1, row)
            dst[:] = src[idx]
        return 0

    def scgr_copy_segments(self, table, n_segments, stream):
        import ctypes as C
        import numpy as np
        _NumpyDouble.calls += 1
        for a in range(n_segments):
            n = table[a].n_floats
            dst = np.ctypeslib.as_array((C.c_float * n).from_address(table[a].dst))
            dst[:] = np.ctypeslib.as_array((C.c_float * n).from_address(table[a].src)) if table[a].src else 0.0
        return 0


@pytest.fixture()
def densify_on_host(reference_renderer, monkeypatch):
    """scgaussian_b200.densify with the C entry points replaced by _NumpyDouble, and a builder of a small CPU model
    through the reference's own GaussianModel / training_setup."""
    import argparse
    import contextlib
    import types
    from scene.gaussian_model import GaussianModel
    from arguments import OptimizationParams
    from scgaussian_b200 import densify

    _NumpyDouble.calls = 0
    monkeypatch.setattr(densify._lib, "load", lambda: _NumpyDouble())
    monkeypatch.setattr(densify, "_require_cuda", lambda device, what: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=0))
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))

    def build(adam_cls, n_ray=9, n_bg=6, K=16):
        g = torch.Generator().manual_seed(21)
        pc = GaussianModel(3)
        P = torch.nn.Parameter
        rnd = lambda *s: torch.randn(*s, generator=g)     # noqa: E731
        pc._rayo, pc._rayd, pc._zval = rnd(n_ray, 3), rnd(n_ray, 3), P(rnd(n_ray, 1))
        pc._features_dc, pc._features_rest = P(rnd(n_ray, 1, 3)), P(rnd(n_ray, K - 1, 3))
        pc._scaling, pc._rotation, pc._opacity = P(rnd(n_ray, 3)), P(rnd(n_ray, 4)), P(rnd(n_ray, 1))
        pc.bg_xyz, pc.bg_features_dc, pc.bg_features_rest = P(rnd(n_bg, 3)), P(rnd(n_bg, 1, 3)), P(rnd(n_bg, K - 1, 3))
        pc.bg_scaling, pc.bg_rotation, pc.bg_opacity = P(rnd(n_bg, 3)), P(rnd(n_bg, 4)), P(rnd(n_bg, 1))
        pc.spatial_lr_scale = 1.0
        pc.training_setup(OptimizationParams(argparse.ArgumentParser()))       # reference :486-512
        if adam_cls is not None:                                                # same groups on the fused optimizer
            pc.optimizer = adam_cls(pc.optimizer.param_groups, lr=0.0, eps=1e-15)
            pc.optimizer_bg = adam_cls(pc.optimizer_bg.param_groups, lr=0.0, eps=1e-15)
        for opt in (pc.optimizer, pc.optimizer_bg):                             # moments as after some steps
            for grp in opt.param_groups:
                p = grp["params"][0]
                if grp["name"] not in ("f_rest", "bg_scaling"):                 # two groups without state (:790, :838)
                    opt.state[p] = {"step": torch.tensor(3.0), "exp_avg": rnd(*p.shape), "exp_avg_sq": rnd(*p.shape).abs()}
        pc.xyz_gradient_accum, pc.denom = rnd(n_ray + n_bg, 1), rnd(n_ray + n_bg, 1).abs()
        pc.max_radii2D = rnd(n_ray + n_bg).abs()
        return pc

    return densify, build


def _same_model(ours, ref, densify, stepped=True):
    attrs = ["_rayo", "_rayd", "xyz_gradient_accum", "denom", "max_radii2D"] + list(densify.GROUP_ATTR.values())
    for a in attrs:
        x, y = getattr(ours, a), getattr(ref, a)
        assert x.shape == y.shape and torch.equal(x.detach(), y.detach()), a
    for o_opt, r_opt in ((ours.optimizer, ref.optimizer), (ours.optimizer_bg, ref.optimizer_bg)):
        for og, rg in zip(o_opt.param_groups, r_opt.param_groups):
            assert og["name"] == rg["name"] and og["lr"] == rg["lr"]
            op, rp = og["params"][0], rg["params"][0]
            assert isinstance(op, torch.nn.Parameter) and op.requires_grad
            assert getattr(ours, densify.GROUP_ATTR[og["name"]]) is op           # the model holds the optimizer's tensor
            assert (op in o_opt.state) == (rp in r_opt.state), og["name"]
            if rp in r_opt.state:
                for k in ("exp_avg", "exp_avg_sq"):
                    assert torch.equal(o_opt.state[op][k], r_opt.state[rp][k]), (og["name"], k)
                assert float(o_opt.state[op]["step"]) == 3.0 or not stepped
        assert len(o_opt.state) == len(r_opt.state)


def test_prune_points_mirror_matches_the_reference_method(densify_on_host):
    """scgaussian_b200.densify.prune_points against the reference's own
    `GaussianModel.prune_points` (scene/gaussian_model.py:795-820), both on the same CPU model built by the reference's
    `training_setup`.  Host logic only (see _NumpyDouble)."""
    from scgaussian_b200 import optim
    densify, build = densify_on_host
    mask = torch.tensor([0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0], dtype=torch.bool)
    ref = build(None)
    ref.prune_points(mask.clone())                                              # the reference's own method
    ours = build(optim.Adam)
    densify.prune_points(ours, mask.clone())
    assert _NumpyDouble.calls == 3                                              # ray set, free set, statistics
    _same_model(ours, ref, densify)


def test_densification_postfix_mirror_matches_the_reference_method(densify_on_host):
    """scgaussian_b200.densify.densification_postfix against the reference's own method
    (scene/gaussian_model.py:844-862 -> cat_tensors_to_optimizer :822-842): the new Gaussians appended to the free set,
    moments extended with zeros, statistics reset -- one launch."""
    from scgaussian_b200 import optim
    densify, build = densify_on_host
    g = torch.Generator().manual_seed(5)
    n_new, K = 5, 16
    new = [torch.randn(n_new, 3, generator=g), torch.randn(n_new, 1, 3, generator=g), torch.randn(n_new, K - 1, 3, generator=g),
           torch.randn(n_new, 1, generator=g), torch.randn(n_new, 3, generator=g), torch.randn(n_new, 4, generator=g)]
    ref = build(None)
    ref.densification_postfix(*[t.clone() for t in new])                        # the reference's own method
    ours = build(optim.Adam)
    densify.densification_postfix(ours, *[t.clone() for t in new])
    assert _NumpyDouble.calls == 1
    _same_model(ours, ref, densify)
    assert ours.bg_xyz.shape == (11, 3) and float(ours.denom.abs().max()) == 0.0 and ours.max_radii2D.shape == (20,)
    # and then the pruning of :915-930 on the grown model
    mask = torch.zeros(20, dtype=torch.bool)
    mask[[10, 16, 19]] = True
    ref.prune_points(mask.clone())
    densify.prune_points(ours, mask.clone())
    _same_model(ours, ref, densify)


@pytest.mark.parametrize("max_screen_size", [None, 20])
def test_densify_and_prune_mirror_matches_the_reference_method(densify_on_host, max_screen_size):
    """scgaussian_b200.densify.densify_and_prune (clone / split selection) against the
    reference's own `GaussianModel.densify_and_prune` (scene/gaussian_model.py:864-931) on the same CPU model and the same
    random stream: identical parameters, moments, ray geometry and statistics afterwards.  Host logic only."""
    from scgaussian_b200 import optim
    densify, build = densify_on_host

    def prepared(adam_cls):
        pc = build(adam_cls, n_ray=40, n_bg=30)
        g = torch.Generator().manual_seed(77)
        P = 70
        pc.percent_dense = 0.01
        with torch.no_grad():
            # scales on both sides of percent_dense * extent, some opacities below the pruning threshold
            pc._scaling.copy_(torch.log(torch.rand(40, 3, generator=g) * 0.04 + 1e-3))
            pc.bg_scaling.copy_(torch.log(torch.rand(30, 3, generator=g) * 0.04 + 1e-3))
            pc._opacity.copy_(torch.randn(40, 1, generator=g) * 3)
            pc.bg_opacity.copy_(torch.randn(30, 1, generator=g) * 3)
        pc.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 0.001
        pc.denom = torch.randint(0, 3, (P, 1), generator=g).float()          # zeros: 0 / 0 -> nan -> 0 (:916)
        pc.max_radii2D = torch.rand(P, generator=g) * 60
        return pc

    ref = prepared(None)
    torch.manual_seed(1234)
    ref.densify_and_prune(0.0002, 0.05, 2.0, max_screen_size)                   # the reference's own method
    ours = prepared(optim.Adam)
    torch.manual_seed(1234)
    densify.densify_and_prune(ours, 0.0002, 0.05, 2.0, max_screen_size)
    _same_model(ours, ref, densify, stepped=False)
    n = ours._zval.shape[0] + ours.bg_xyz.shape[0]
    assert ours._zval.shape[0] == 40 and n != 70                                # ray-based Gaussians are never pruned; the free set changed
    assert ours.max_radii2D.shape == (n,)


def test_build_rotation_matches_the_reference(reference_renderer, monkeypatch):
    from utils.general_utils import build_rotation as ref_build
    from scgaussian_b200.densify import build_rotation
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))
    q = torch.randn(257, 4, generator=torch.Generator().manual_seed(3))
    assert torch.equal(build_rotation(q), ref_build(q))


# ---------------------------------------------------------------------------------------------------------------
# On the GPU: the reference's own render() + GaussianModel accessors, unmodified, against the real kernels
# ---------------------------------------------------------------------------------------------------------------
def _reference_model(G, case, n_ray, dev, sh_degree=3):
    """A reference GaussianModel (scene/gaussian_model.py) holding `case` as a hybrid of n_ray ray-based Gaussians
    (position = rayo + rayd * zval, :124) and free ones (bg_*), raw parameters such that the reference's OWN activations
    (:105-152) reproduce the case's scales / rotations / opacities."""
    from scene.gaussian_model import GaussianModel
    pc = GaussianModel(sh_degree)
    pc.active_sh_degree = case["sh_degree"]
    m = case["means3D"]
    g = torch.Generator().manual_seed(9)
    rayo = torch.randn(n_ray, 3, generator=g) * 0.1
    d = m[:n_ray] - rayo
    z = d.norm(dim=1, keepdim=True)
    par = lambda t: torch.nn.Parameter(t.clone().to(dev).contiguous())       # noqa: E731
    pc._rayo, pc._rayd, pc._zval = rayo.to(dev), (d / z).to(dev), par(z)
    sc, ro, op, sh = case["scales"].log(), case["rotations"] * 1.7, torch.logit(case["opacities"]), case["shs"]
    pc._scaling, pc._rotation, pc._opacity = par(sc[:n_ray]), par(ro[:n_ray]), par(op[:n_ray])
    pc._features_dc, pc._features_rest = par(sh[:n_ray, :1]), par(sh[:n_ray, 1:])
    if n_ray < case["P"]:
        pc.bg_xyz = par(m[n_ray:])
        pc.bg_scaling, pc.bg_rotation, pc.bg_opacity = par(sc[n_ray:]), par(ro[n_ray:]), par(op[n_ray:])
        pc.bg_features_dc, pc.bg_features_rest = par(sh[n_ray:, :1]), par(sh[n_ray:, 1:])
    return pc


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["default", "convert_SHs_python", "compute_cov3D_python"])
def test_reference_render_runs_on_the_gpu_against_libscgr(reference_renderer, variant):
    """SURVEY.md section 8a rows a1 + a2: reference gaussian_renderer/__init__.py:20-118 and the GaussianModel accessors
    of scene/gaussian_model.py:105-152, imported unmodified, executed on the GPU with `diff_gaussian_rasterization`
    resolving to this repo -- images and every parameter gradient against the CPU oracle fed with the activations the
    reference computed, `viewspace_points.grad` against the oracle's NDC-scaled screen gradient, and the two
    reference-owned switches (`convert_SHs_python`, `compute_cov3D_python`) against the default path."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import math
    import numpy as np
    from oracle import torch_oracle as O
    from tests import util
    G = reference_renderer
    dev = torch.device("cuda:0")
    P, W, H = 4000, 200, 152
    case = util.make_case(P, W, H, sh_degree=3, scale_median=0.05, bg=(0.1, 0.2, 0.3), w2c=O.yaw_w2c(6.0), seed=11)
    n_ray = P if variant == "compute_cov3D_python" else 2500     # (the reference's get_covariance ignores the bg set, :152)
    pc = _reference_model(G, case, n_ray, dev)

    class Cam:                                   # what reference scene/cameras.py:54-63 exposes
        FoVx, FoVy = 2 * math.atan(case["tanfovx"]), 2 * math.atan(case["tanfovy"])
        image_height, image_width = H, W
        world_view_transform = case["viewmatrix"].to(dev)
        full_proj_transform = case["projmatrix"].to(dev)
        camera_center = case["campos"].to(dev)

    class Pipe:
        convert_SHs_python = variant == "convert_SHs_python"
        compute_cov3D_python = variant == "compute_cov3D_python"
        debug = False

    out = G.render(Cam(), pc, Pipe(), case["bg"].to(dev))
    assert set(out) >= {"render", "rendered_depth", "rendered_alpha", "viewspace_points", "visibility_filter", "radii"}
    gC, gD, gA = O.synth_upstream_grads(W, H)
    loss = (out["render"] * gC.to(dev)).sum() + (out["rendered_depth"] * gD.to(dev)).sum() + (out["rendered_alpha"] * gA.to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    # the oracle sees what the reference's accessors produced
    with torch.no_grad():
        act = dict(case, means3D=pc.get_xyz.cpu(), scales=pc.get_scaling.cpu(), rotations=pc.get_rotation.cpu(),
                   opacities=pc.get_opacity.cpu(), shs=pc.get_features.cpu())
    co, (c2, r2, d2, a2), g2 = util.run_c_oracle(act, "f32", grads=(gC, gD, gA))
    flips = util.flip_sets(co)
    assert out["radii"].dtype == torch.int32 and out["render"].shape == (3, H, W)
    util.assert_radii_match("radii", out["radii"].cpu().numpy(), r2, flips)
    assert torch.equal(out["visibility_filter"], out["radii"] > 0)
    util.assert_image_close("render", out["render"].detach().cpu().numpy(), c2, flips)
    util.assert_image_close("rendered_depth", out["rendered_depth"].detach().cpu().numpy(), d2, flips)
    util.assert_image_close("rendered_alpha", out["rendered_alpha"].detach().cpu().numpy(), a2, flips)
    # screenspace_points.grad: the hook of reference lines 28-32, in NDC units, z = 0
    vg = out["viewspace_points"].grad
    assert vg is not None and float(vg[:, 2].abs().max()) == 0.0
    util.assert_grad_close("viewspace_points.grad", vg.cpu().numpy(), g2["means2D"], flips)
    # parameter gradients through the reference's own activations: chain the oracle's gradients through the same torch ops
    leaves = {k: act[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    ref = _reference_model(G, case, n_ray, torch.device("cpu"))
    pairs = [(ref.get_xyz, "means3D"), (ref.get_scaling, "scales"), (ref.get_rotation, "rotations"),
             (ref.get_opacity, "opacities"), (ref.get_features, "shs")]
    sur = sum((t * torch.from_numpy(np.asarray(g2[k], dtype=np.float32)).reshape(t.shape)).sum() for t, k in pairs)
    sur.backward()
    del leaves
    names = ["_zval", "_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest"]
    if n_ray < P:
        names += ["bg_xyz", "bg_scaling", "bg_rotation", "bg_opacity", "bg_features_dc", "bg_features_rest"]
    for n in names:
        got, want = getattr(pc, n).grad, getattr(ref, n).grad
        assert got is not None and want is not None, n
        sl = slice(n_ray, None) if n.startswith("bg_") else slice(0, n_ray)
        part = dict(flips, **{k: flips[k][sl] for k in ("gauss_flag", "gauss_margin", "gauss_own")})
        util.assert_grad_close(n, got.cpu().numpy(), want.numpy(), part)


@pytest.mark.gpu
def test_densify_and_prune_on_the_gpu_matches_the_reference_method(reference_renderer):
    """the reference's own `GaussianModel.densify_and_prune` (scene/gaussian_model.py:864-931,
    unmodified, its two torch optimizers) against scgaussian_b200.densify.densify_and_prune on the fused optimizer -- the
    real gather / append kernels -- on the same CUDA model and the same CUDA random stream: identical model afterwards."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import argparse
    from arguments import OptimizationParams
    from scgaussian_b200 import densify, optim
    from tests import util
    G = reference_renderer
    dev = torch.device("cuda:0")
    P, n_ray = 30_000, 18_000
    case = util.make_case(P, 320, 200, sh_degree=3, scale_median=0.02, seed=4)

    def prepared(fused):
        pc = _reference_model(G, case, n_ray, dev)
        pc.spatial_lr_scale = 1.0
        pc.training_setup(OptimizationParams(argparse.ArgumentParser()))       # reference :486-512
        if fused:
            pc.optimizer = optim.Adam(pc.optimizer.param_groups, lr=0.0, eps=1e-15)
            pc.optimizer_bg = optim.Adam(pc.optimizer_bg.param_groups, lr=0.0, eps=1e-15)
        g = torch.Generator().manual_seed(8)
        for opt in (pc.optimizer, pc.optimizer_bg):                             # moments as after some steps
            for grp in opt.param_groups:
                p = grp["params"][0]
                if grp["name"] not in ("f_rest", "bg_scaling"):
                    opt.state[p] = {"step": torch.tensor(3.0), "exp_avg": torch.randn(p.shape, generator=g).to(dev),
                                    "exp_avg_sq": torch.randn(p.shape, generator=g).abs().to(dev)}
        pc.xyz_gradient_accum = (torch.rand(P, 1, generator=g) * 0.001).to(dev)
        pc.denom = torch.randint(0, 3, (P, 1), generator=g).float().to(dev)
        pc.max_radii2D = (torch.rand(P, generator=g) * 60).to(dev)
        return pc

    ref = prepared(False)
    torch.manual_seed(99)
    ref.densify_and_prune(0.0002, 0.05, 2.0, 20)                                # the reference's own method
    ours = prepared(True)
    torch.manual_seed(99)
    densify.densify_and_prune(ours, 0.0002, 0.05, 2.0, 20)
    torch.cuda.synchronize()
    _same_model(ours, ref, densify, stepped=False)
    assert ours._zval.shape[0] == n_ray and ours.bg_xyz.shape[0] != P - n_ray
