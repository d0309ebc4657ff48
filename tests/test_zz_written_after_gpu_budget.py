"""GPU tests written AFTER this round's GPU budget was spent: they have not run on a B200 yet.  The kernels they
cover were checked by host emulation (every thread of every block executed on the CPU, barriers included) and the
host logic against the reference's own methods (tests/test_reference_callsite.py).  The file sorts last so that,
under `pytest -x`, a surprise here cannot mask the verified suite; fold the tests into test_model.py /
test_parity_gpu.py once they have been seen green on the GPU."""
import pytest
import torch

from tests import util
from tests.test_model import _Stats
from tests.test_parity_gpu import dev, settings_for  # noqa: F401  (dev: module-scoped fixture)

pytestmark = pytest.mark.gpu


def test_densification_postfix_matches_torch_cat():
    """scgaussian_b200.densify.densification_postfix (reference scene/gaussian_model.py:822-862) on the GPU: parameters
    = torch.cat(old, new), moments = torch.cat(old, zeros), statistics reset -- bit for bit, in one launch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from scgaussian_b200 import _lib, densify, optim
    gen = torch.Generator().manual_seed(17)
    n_ray, n_bg, n_new, K = 3001, 4099, 1234, 16

    def rnd(*s):
        return torch.randn(*s, generator=gen).cuda()
    pc = _Stats()
    par = torch.nn.Parameter
    pc._zval = par(rnd(n_ray, 1))
    pc.bg_xyz, pc.bg_features_dc, pc.bg_features_rest = par(rnd(n_bg, 3)), par(rnd(n_bg, 1, 3)), par(rnd(n_bg, K - 1, 3))
    pc.bg_scaling, pc.bg_rotation, pc.bg_opacity = par(rnd(n_bg, 3)), par(rnd(n_bg, 4)), par(rnd(n_bg, 1))
    inv = {v: k for k, v in densify.GROUP_ATTR.items()}
    free = [a for a in densify.GROUP_ATTR.values() if a.startswith("bg_")]
    pc.optimizer_bg = optim.Adam([{"params": [getattr(pc, a)], "lr": 1e-3, "name": inv[a]} for a in free], lr=0.0, eps=1e-15)
    for a in free:
        if a != "bg_opacity":                       # one group that has never been stepped: no state to extend
            getattr(pc, a).grad = rnd(*getattr(pc, a).shape)
    pc.optimizer_bg.step()
    pc.xyz_gradient_accum, pc.denom, pc.max_radii2D = rnd(n_ray + n_bg, 1), rnd(n_ray + n_bg, 1), rnd(n_ray + n_bg)
    new = {"bg_xyz": rnd(n_new, 3), "bg_features_dc": rnd(n_new, 1, 3), "bg_features_rest": rnd(n_new, K - 1, 3),
           "bg_opacity": rnd(n_new, 1), "bg_scaling": rnd(n_new, 3), "bg_rotation": rnd(n_new, 4)}
    want_p = {a: torch.cat((getattr(pc, a).detach(), new[a])) for a in free}
    want_m = {a: tuple(torch.cat((pc.optimizer_bg.state[getattr(pc, a)][k], torch.zeros_like(new[a])))
                       for k in ("exp_avg", "exp_avg_sq")) for a in free if a != "bg_opacity"}
    lib = _lib.load()
    before = lib.scgr_kernel_launch_count()
    densify.densification_postfix(pc, new["bg_xyz"], new["bg_features_dc"], new["bg_features_rest"], new["bg_opacity"],
                                  new["bg_scaling"], new["bg_rotation"])
    assert lib.scgr_kernel_launch_count() - before == 1
    P = n_ray + n_bg + n_new
    for a in free:
        p = getattr(pc, a)
        assert isinstance(p, torch.nn.Parameter) and p.requires_grad and torch.equal(p.detach(), want_p[a]), a
        grp = [g for g in pc.optimizer_bg.param_groups if g["name"] == inv[a]][0]
        assert grp["params"][0] is p
        if a == "bg_opacity":
            assert p not in pc.optimizer_bg.state
        else:
            st = pc.optimizer_bg.state[p]
            assert torch.equal(st["exp_avg"], want_m[a][0]) and torch.equal(st["exp_avg_sq"], want_m[a][1]), a
            assert float(st["step"]) == 1.0
    assert len(pc.optimizer_bg.state) == 5
    for t, shape in ((pc.xyz_gradient_accum, (P, 1)), (pc.denom, (P, 1)), (pc.max_radii2D, (P,))):
        assert tuple(t.shape) == shape and float(t.abs().max()) == 0.0
    # the grown model keeps training
    for a in free:
        getattr(pc, a).grad = torch.ones_like(getattr(pc, a))
    pc.optimizer_bg.step()
    assert float(pc.optimizer_bg.state[pc.bg_xyz]["step"]) == 2.0 and float(pc.optimizer_bg.state[pc.bg_opacity]["step"]) == 1.0
    # nothing to append
    empty = {a: new[a][:0] for a in free}
    densify.densification_postfix(pc, empty["bg_xyz"], empty["bg_features_dc"], empty["bg_features_rest"], empty["bg_opacity"],
                                  empty["bg_scaling"], empty["bg_rotation"])
    assert pc.bg_xyz.shape == (n_bg + n_new, 3)


def test_overflow_flag_describes_the_last_emission(dev, monkeypatch):
    """A fused call whose pre-sized buffer is too small raises the device overflow flag and returns
    SCGR_NEED_CAPACITY; the completing scgr_forward_render (stage 1 kept) must leave {R, 0} behind, with the same
    tile lists as a forward that never overflowed."""
    from scgaussian_b200 import rasterizer as R
    case = util.make_case(4000, 160, 120, scale_median=0.05)      # test_binning_modes_agree's scene: R >> 4096 + 20
    s = settings_for(case, dev)
    t = {k: case[k].to(dev).contiguous() for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    args = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None, s)
    monkeypatch.setattr(R, "_BINNING_MODE", "sync")
    ref = R.rasterize_forward_raw(*args)
    dv_ref = R.debug_views(ref[4], case["P"], s)
    monkeypatch.setattr(R, "_BINNING_MODE", "fused")
    R._capacity_hint[dev.index] = 16
    out = R.rasterize_forward_raw(*args)
    torch.cuda.synchronize()
    assert out[4].capacity == out[4].num_rendered == ref[4].num_rendered
    dv = R.debug_views(out[4], case["P"], s)
    assert int(dv["status"][0]) == ref[4].num_rendered and int(dv["status"][1]) == 0
    assert torch.equal(dv["point_list"], dv_ref["point_list"]) and torch.equal(dv["ranges"], dv_ref["ranges"])
    assert torch.equal(out[0], ref[0])
