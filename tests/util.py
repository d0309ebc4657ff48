"""Shared helpers of the parity tests: scene construction, oracle front-ends, tolerant compares."""
from __future__ import annotations

import math
import os

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
# element beyond rtol is therefore accepted ONLY when the oracle proves that a decision it depends on lies inside the
# uncertainty band of its threshold, the band following from a stated fp32 error model (ERROR_MODEL below;
# oracle/scg_oracle.c: scgo_margins): the pixel is flip-prone, or the Gaussian (nearly) contributes to a flip-prone
# pixel.  Such elements must still lie within FLIP_BOUND norm-wise -- one flipped contribution is worth at most
# alpha at the edge of the 3-sigma rect, 0.011 x opacity, of the value range.  Without an oracle state
# (`flips=None`) nothing is excused.  Every compare returns what it measured, including `model_scale_needed`: the
# factor by which the error model would have had to be scaled to excuse the worst offending element (< 1 = inside).
# SCGR_PARITY_CALIBRATE=1 turns the "unexplained element" failure into a record (to size the model from a first run).
# ---------------------------------------------------------------------------------------------------------------
ERROR_MODEL = dict(base_err=2e-6,      # relative: exp() + the products around it (hardware exp2: 2 ulp)
                   conic_err=2e-6,     # relative: conic coefficients, hence power
                   pos_ulps=2.5)       # projected mean: 2.5 ulp of the largest pixel coordinate (2^-23 max(W, H) px each);
#                                        measured on B200 over every parity case: <= 1.95 (profiles/r02_parity.jsonl)
FLIP_BOUND = 1.2e-2
IMAGE_FLOOR = 0.01       # x RMS of the image
GRAD_FLOOR = 1.0         # x RMS of the non-zero gradient entries
CALIBRATE = os.environ.get("SCGR_PARITY_CALIBRATE") == "1"


def flip_sets(co, **model):
    """Flip-prone pixels / flip-affected Gaussians of the view the C oracle `co` last rendered."""
    kw = dict(ERROR_MODEL)
    kw.update(model)
    return co.margins(**kw)


def _elementwise(got, want, floor_frac, nonzero_rms):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    sel = want[want != 0] if nonzero_rms else want
    rms = float(np.sqrt(np.mean(sel * sel))) if sel.size else 0.0
    diff = np.abs(got - want)
    return diff / (np.abs(want) + floor_frac * rms + 1e-300), diff / (np.abs(want).max() + 1e-30)


def _judge(name, err, nw, margin, rtol):
    """margin: per-element margin / uncertainty of the nearest decision the element depends on (None: nothing is
    excusable).  An element is excusable when margin < 1."""
    bad = err > rtol
    n_bad = int(bad.sum())
    excusable = np.zeros(err.shape, bool) if margin is None else margin < 1.0
    stats = dict(max_err=float(err.max()) if err.size else 0.0, max_normwise=float(nw.max()) if nw.size else 0.0,
                 n=int(err.size), n_beyond_rtol=n_bad, excusable_frac=float(excusable.mean()) if excusable.size else 0.0,
                 max_err_unexcusable=float(err[~excusable].max()) if (~excusable).any() else 0.0)
    if n_bad:
        stats["model_scale_needed"] = float(margin[bad].max()) if margin is not None else float("inf")
        stats["max_normwise_beyond_rtol"] = float(nw[bad].max())
        rogue = bad & ~excusable
        if CALIBRATE:
            stats["n_unexplained"] = int(rogue.sum())
            return stats
        assert not rogue.any(), (f"{name}: {int(rogue.sum())} element(s) beyond rtol={rtol} that no decision inside its "
                                 f"uncertainty band explains (max err {err[rogue].max():.3e}; {n_bad} beyond rtol in all; "
                                 f"error model would need x{stats['model_scale_needed']:.3g})")
        assert nw[bad].max() <= FLIP_BOUND, f"{name}: a flip-excused element is off by {nw[bad].max():.3e} > {FLIP_BOUND} norm-wise"
    return stats


def assert_image_close(name, got, want, flips=None, rtol=RTOL_IMAGE, normwise=False):
    """[C,H,W] images agree element-wise within rtol; elements beyond it must be flip-prone pixels (see above).
    normwise=True judges max|a - b| / max|b| per element instead (north_star's literal "rel"; used for the end-to-end
    compare, where the two preprocess implementations legitimately differ by ulps of the pixel coordinates -- at 4K
    that alone moves alpha by ~1e-4 relative, element-wise; the element-wise statement is made on identical 2D state).
    Returns the measured figures (recorded in profiles/r02_parity.jsonl by the GPU tests)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    err, nw = _elementwise(got, want, IMAGE_FLOOR, False)
    margin = None if flips is None else np.broadcast_to(flips["pix_margin"][None].astype(np.float64), got.shape)
    st = _judge(name, nw if normwise else err, nw, margin, rtol)
    st["metric"] = "normwise" if normwise else "elementwise"
    st["max_err_elementwise"] = float(err.max()) if err.size else 0.0
    return st


def assert_grad_close(name, got, want, flips=None, rtol=RTOL_GRAD, normwise=False):
    """[P,...] per-Gaussian gradients agree element-wise within rtol; rows beyond it must belong to flip-affected
    Gaussians (see above).  normwise: as in assert_image_close."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if got.size == 0:
        return dict(max_err=0.0, max_normwise=0.0, n=0, n_beyond_rtol=0)
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    err, nw = _elementwise(got, want, GRAD_FLOOR, True)
    margin = None
    if flips is not None:
        gm = np.where(flips["gauss_flag"], np.minimum(flips["gauss_margin"], 0.999), flips["gauss_margin"]).astype(np.float64)
        margin = np.broadcast_to(gm.reshape((-1,) + (1,) * (got.ndim - 1)), got.shape)
    st = _judge(name, nw if normwise else err, nw, margin, rtol)
    st["metric"] = "normwise" if normwise else "elementwise"
    st["max_err_elementwise"] = float(err.max())
    if flips is not None and st["n_beyond_rtol"]:
        rows = ((nw if normwise else err) > rtol).reshape(len(err), -1).any(1)
        st["rows_beyond_rtol"] = int(rows.sum())
        st["rows_beyond_rtol_own_decision"] = int((rows & flips["gauss_own"]).sum())
    return st


def assert_radii_match(name, got, want, flips=None):
    """radii are integers: bit-exact, except where 3 sqrt(lambda) sits within the model's geometric uncertainty of an
    integer (ceil of a last-bit difference) -- then they may differ by one.  Returns the number of such mismatches."""
    got, want = np.asarray(got).astype(np.int64), np.asarray(want).astype(np.int64)
    mis = got != want
    n = int(mis.sum())
    if n:
        assert flips is not None, f"{name}: {n} radii differ"
        assert np.abs(got - want)[mis].max() <= 1, f"{name}: radii differ by more than one"
        gm, lim = flips["geom_margin"][mis], flips["model"]["geom_err_px"]
        if not CALIBRATE:
            assert (gm < lim).all(), (f"{name}: {n} radii differ, {int((gm >= lim).sum())} of them away from a rounding "
                                      f"boundary (margin up to {gm.max():.3e} px, model {lim:.3e} px)")
    return n


# ---- staged parity: the oracle's binning + blend + backward on the 2D state the CUDA preprocess produced ----
STAGED_MODEL = dict(base_err=2e-6, conic_err=5e-7, pos_ulps=0.0)     # identical means / radii; conic re-scaled once


def records_override(record, radii):
    """libscgr's packed per-Gaussian record [P,12] (include/scgr.h: ScgrDebugViews.record) -> the override arrays of
    COracle.forward: means2D, conic (un-scaled from the render-ready log2 form), rgb, depth, radii, clamp flags."""
    rec = np.ascontiguousarray(record, dtype=np.float32)
    radii = np.ascontiguousarray(radii, dtype=np.int32)
    log2e = 1.4426950408889634
    r64 = rec.astype(np.float64)
    conic = np.stack([r64[:, 2] / (-0.5 * log2e), r64[:, 3] / (-log2e), r64[:, 4] / (-0.5 * log2e)], 1)
    bits = rec[:, 11].copy().view(np.uint32)
    flags = (bits >> 28).astype(np.uint8)
    clamped = np.stack([flags & 1, (flags >> 1) & 1, (flags >> 2) & 1], 1).astype(np.uint8)
    vis = radii > 0
    z = lambda a: np.where(vis.reshape((-1,) + (1,) * (a.ndim - 1)), a, 0)
    return dict(xy=z(r64[:, 0:2]), conic=z(conic), rgb=z(r64[:, 8:11]), depth=z(r64[:, 6]), radii=radii, clamped=z(clamped))


def run_c_oracle_staged(case, record, radii, precision="f32", grads=None, colors_precomp=None, cov3D_precomp=None):
    """run_c_oracle with the 2D state forced to the CUDA preprocess's (see records_override)."""
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
                     scale_modifier=case["scale_modifier"], override=records_override(record, radii), **kw)
    g = co.backward(*[x.numpy() for x in grads]) if grads is not None else None
    return co, out, g


def preprocess_errors(record, radii, co):
    """The CUDA preprocess's per-Gaussian values against the oracle's own (A.1-A.5), over the Gaussians both keep."""
    st = co.state()
    rec = np.asarray(record, dtype=np.float64)
    radii = np.asarray(radii)
    both = (radii > 0) & (st["tiles_touched"] > 0)
    log2e = 1.4426950408889634
    conic = np.stack([rec[:, 2] / (-0.5 * log2e), rec[:, 3] / (-log2e), rec[:, 4] / (-0.5 * log2e)], 1)
    W, H = co._dims[2], co._dims[3]
    ulp = 2.0 ** -23 * max(W, H)
    cs = np.abs(st["conic"][both]).max(1, keepdims=True) + 1e-30        # per-Gaussian scale of its conic
    return dict(n_both=int(both.sum()),
                xy_px=float(np.abs(rec[both][:, 0:2] - st["means2D"][both]).max()) if both.any() else 0.0,
                xy_ulps_of_max_coord=float(np.abs(rec[both][:, 0:2] - st["means2D"][both]).max() / ulp) if both.any() else 0.0,
                conic_rel=float((np.abs(conic[both] - st["conic"][both]) / cs).max()) if both.any() else 0.0,
                rgb_abs=float(np.abs(rec[both][:, 8:11] - st["rgb"][both]).max()) if both.any() else 0.0,
                depth_rel=float((np.abs(rec[both][:, 6] - st["depths"][both]) / st["depths"][both]).max()) if both.any() else 0.0)
