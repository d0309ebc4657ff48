"""Shared helpers of the parity tests: scene construction, oracle front-ends, tolerant compares."""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle import torch_oracle as O
from oracle.c_oracle import COracle

# north_star tolerances
RTOL_IMAGE = 1e-4      # RGB / depth / alpha
RTOL_GRAD = 1e-3       # every returned gradient


def make_case(P, W, H, sh_degree=3, scale_median=0.05, seed=0, w2c=None, bg=(0.0, 0.0, 0.0),
              scale_modifier=1.0, max_sh_degree=3, fovx_deg=60.0, z_shift=0.0):
    cam = O.make_camera(W, H, fovx_deg=fovx_deg, w2c=w2c)
    sc = O.synth_scene(P, W, H, sh_degree=sh_degree, max_sh_degree=max_sh_degree,
                       scale_median=scale_median, seed=seed, fovx_deg=fovx_deg)
    if z_shift:
        sc["means3D"][:, 2] += z_shift
    return dict(P=P, W=W, H=H, sh_degree=sh_degree, scale_modifier=scale_modifier,
                bg=torch.tensor(bg, dtype=torch.float32), **cam, **sc)


def oracle_settings(case, dtype=torch.float32):
    return O.Settings(case["H"], case["W"], case["tanfovx"], case["tanfovy"], case["bg"].to(dtype),
                      case["scale_modifier"], case["viewmatrix"].to(dtype), case["projmatrix"].to(dtype),
                      case["sh_degree"], case["campos"].to(dtype))


def run_c_oracle(case, precision="f32", grads=None, colors_precomp=None, cov3D_precomp=None):
    co = COracle(precision)
    kw = {}
    if colors_precomp is not None:
        kw["colors_precomp"] = colors_precomp.numpy()
    else:
        kw["shs"] = case["shs"].numpy()
    if cov3D_precomp is not None:
        kw["cov3D_precomp"] = cov3D_precomp.numpy()
    else:
        kw["scales"] = case["scales"].numpy()
        kw["rotations"] = case["rotations"].numpy()
    out = co.forward(means3D=case["means3D"].numpy(), opacities=case["opacities"].numpy(), W=case["W"],
                     H=case["H"], tanfovx=case["tanfovx"], tanfovy=case["tanfovy"], bg=case["bg"].numpy(),
                     viewmatrix=case["viewmatrix"].numpy(), projmatrix=case["projmatrix"].numpy(),
                     campos=case["campos"].numpy(), sh_degree=case["sh_degree"],
                     scale_modifier=case["scale_modifier"], **kw)
    g = None
    if grads is not None:
        g = co.backward(*[x.numpy() for x in grads])
    return co, out, g


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error; reported next to the element-wise figure)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# ---------------------------------------------------------------------------------------------------------------
# Parity compares.  Tolerance (north_star): 1e-4 rel on RGB / depth / alpha, 1e-3 rel on every returned gradient,
# read ELEMENT-WISE:  |got - want| <= rtol * (|want| + floor)  with a small floor that keeps elements near zero
# meaningful -- 1 % of the array's RMS for images, the RMS of the non-zero entries for gradients (whose entries are
# sums of signed per-pixel terms: an entry that cancels to ~0 carries the rounding of terms of typical size).
#
# The rasterizer also takes DISCRETE decisions (alpha >= 1/255, T (1 - alpha) >= 1e-4, power <= 0, radius = ceil(..),
# tile rect = trunc(..)) that a last-bit difference can flip; a flip moves a pixel by up to ~alpha = 4e-3 -- no
# tolerance on the values can absorb it, the f32 and f64 builds of the oracle differ from EACH OTHER this way.  An
# element beyond rtol is therefore accepted ONLY when the oracle proves that a decision it depends on sits within
# EPS of its threshold (oracle/scg_oracle.c: scgo_margins): the pixel is flip-prone, or the Gaussian (nearly)
# contributes to a flip-prone pixel.  Such elements must still lie within FLIP_BOUND norm-wise -- one flipped
# contribution is worth at most alpha at the edge of the 3-sigma rect, 0.011 x opacity, of the value range.  Without
# an oracle state (`flips=None`) nothing is excused.
# ---------------------------------------------------------------------------------------------------------------
EPS = dict(eps_alpha=5e-5, eps_T=5e-5, eps_power=1e-6, eps_geom=1e-4)    # relative, relative, absolute, pixels
FLIP_BOUND = 1.2e-2
IMAGE_FLOOR = 0.01       # x RMS of the image
GRAD_FLOOR = 1.0         # x RMS of the non-zero gradient entries


def flip_sets(co, **eps):
    """Flip-prone pixels / flip-affected Gaussians of the view the C oracle `co` last rendered."""
    kw = dict(EPS)
    kw.update(eps)
    m = co.margins(**kw)
    m["eps"] = kw
    return m


def _elementwise(got, want, floor_frac, nonzero_rms):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    sel = want[want != 0] if nonzero_rms else want
    rms = float(np.sqrt(np.mean(sel * sel))) if sel.size else 0.0
    diff = np.abs(got - want)
    return diff / (np.abs(want) + floor_frac * rms + 1e-300), diff / (np.abs(want).max() + 1e-30)


def _judge(name, err, nw, excusable, rtol):
    bad = err > rtol
    n_bad = int(bad.sum())
    stats = dict(max_err=float(err.max()) if err.size else 0.0, max_normwise=float(nw.max()) if nw.size else 0.0,
                 n=int(err.size), n_beyond_rtol=n_bad, excusable_frac=float(excusable.mean()) if excusable.size else 0.0,
                 max_err_unexcusable=float(err[~excusable].max()) if (~excusable).any() else 0.0)
    if n_bad:
        rogue = bad & ~excusable
        assert not rogue.any(), (f"{name}: {int(rogue.sum())} element(s) beyond rtol={rtol} that no near-threshold "
                                 f"decision explains (max {err[rogue].max():.3e}; {n_bad} beyond rtol in all)")
        assert nw[bad].max() <= FLIP_BOUND, f"{name}: a flip-excused element is off by {nw[bad].max():.3e} > {FLIP_BOUND} norm-wise"
        stats["max_normwise_excused"] = float(nw[bad].max())
    return stats


def assert_image_close(name, got, want, flips=None, rtol=RTOL_IMAGE):
    """[C,H,W] images agree element-wise within rtol; elements beyond it must be flip-prone pixels (see above).
    Returns the measured figures (recorded in profiles/r02_parity.jsonl by the GPU tests)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    err, nw = _elementwise(got, want, IMAGE_FLOOR, False)
    exc = np.zeros(got.shape, bool) if flips is None else np.broadcast_to(flips["pix_flag"][None], got.shape)
    return _judge(name, err, nw, exc, rtol)


def assert_grad_close(name, got, want, flips=None, rtol=RTOL_GRAD):
    """[P,...] per-Gaussian gradients agree element-wise within rtol; rows beyond it must belong to flip-affected
    Gaussians (see above)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if got.size == 0:
        return dict(max_err=0.0, max_normwise=0.0, n=0, n_beyond_rtol=0)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    err, nw = _elementwise(got, want, GRAD_FLOOR, True)
    exc = np.zeros(got.shape, bool)
    if flips is not None:
        exc = np.broadcast_to(flips["gauss_flag"].reshape((-1,) + (1,) * (got.ndim - 1)), got.shape)
    return _judge(name, err, nw, exc, rtol)


def assert_radii_match(name, got, want, flips=None):
    """radii are integers: bit-exact, except where 3 sqrt(lambda) sits within eps_geom of an integer (ceil of a
    last-bit difference) -- then they may differ by one.  Returns the number of such mismatches."""
    got, want = np.asarray(got).astype(np.int64), np.asarray(want).astype(np.int64)
    mis = got != want
    n = int(mis.sum())
    if n:
        assert flips is not None, f"{name}: {n} radii differ"
        assert np.abs(got - want)[mis].max() <= 1, f"{name}: radii differ by more than one"
        gm = flips["geom_margin"][mis]
        assert (gm < flips["eps"]["eps_geom"]).all(), \
            f"{name}: {n} radii differ, {int((gm >= flips['eps']['eps_geom']).sum())} of them away from a rounding boundary (margin up to {gm.max():.3e} px)"
    return n
