"""CPU tests of the drop-in boundary: libscgr.so loads without a GPU, exports every symbol that
include/scgr.h declares, sizes its scratch sanely and reports argument errors through the C ABI
(status + scgr_last_error) -- no compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

from scgaussian_b200 import _lib
from scgaussian_b200._lib import ScgrGaussians, ScgrGrads, ScgrView

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "scgr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scgr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"libscgr.so does not export {s}"
        assert s in _lib.SYMBOLS, f"python binding misses {s}"
    assert sorted(_lib.SYMBOLS) == syms


def test_version_and_scratch_sizes():
    lib = _lib.load()
    assert lib.scgr_version() == 200
    g1, g2 = lib.scgr_geometry_bytes(1000), lib.scgr_geometry_bytes(1_000_000)
    assert 0 < g1 < g2 and g2 % 256 == 0
    # per-Gaussian footprint stays well under the reference's own geomBuffer + our gradient staging
    assert g2 / 1e6 < 160
    b = lib.scgr_binning_bytes(1_000_000, 1920, 1080, 10_000_000)
    assert 16 * 10_000_000 <= b < 20 * 10_000_000          # two ping-pong (key, value) arrays
    assert lib.scgr_image_bytes(1920, 1080) >= 8 * 1920 * 1080
    assert lib.scgr_geometry_bytes(0) > 0 and lib.scgr_binning_bytes(0, 0, 0, 0) > 0


def _view(**kw):
    d = dict(image_height=32, image_width=32, tanfovx=0.5, tanfovy=0.5, bg=1, scale_modifier=1.0,
             viewmatrix=1, projmatrix=1, sh_degree=0, campos=1, prefiltered=0, debug=0)
    d.update(kw)
    return ScgrView(**d)


def test_argument_errors_come_back_through_the_c_abi():
    lib = _lib.load()
    v = _view()
    # both colour sources given
    g = ScgrGaussians(P=4, sh_coeffs=1, means3D=256, opacities=256, shs=256, colors_precomp=256,
                      scales=256, rotations=256, cov3D_precomp=None)
    rc = lib.scgr_forward_geometry(C.byref(v), C.byref(g), 256, 256, None, None)
    assert rc != 0 and b"exactly one of shs / colors_precomp" in lib.scgr_last_error()
    # neither covariance source
    g = ScgrGaussians(P=4, sh_coeffs=1, means3D=256, opacities=256, shs=256, colors_precomp=None,
                      scales=None, rotations=None, cov3D_precomp=None)
    rc = lib.scgr_forward_geometry(C.byref(v), C.byref(g), 256, 256, None, None)
    assert rc != 0 and b"cov3D_precomp" in lib.scgr_last_error()
    # sh degree out of range / too few coefficients
    g = ScgrGaussians(P=4, sh_coeffs=4, means3D=256, opacities=256, shs=256, colors_precomp=None,
                      scales=256, rotations=256, cov3D_precomp=None)
    rc = lib.scgr_forward_geometry(C.byref(_view(sh_degree=3)), C.byref(g), 256, 256, None, None)
    assert rc != 0 and b"too few coefficients" in lib.scgr_last_error()
    # unaligned scratch
    g = ScgrGaussians(P=4, sh_coeffs=1, means3D=256, opacities=256, shs=256, colors_precomp=None,
                      scales=256, rotations=256, cov3D_precomp=None)
    rc = lib.scgr_forward_geometry(C.byref(v), C.byref(g), 260, 256, None, None)
    assert rc != 0 and b"256-byte aligned" in lib.scgr_last_error()
    # backward: gradient set must mirror the inputs
    gr = ScgrGrads(256, 256, None, None, 256, 256, 256, None)
    rc = lib.scgr_backward(C.byref(v), C.byref(g), 256, 256, 0, 256, 256, 256, 256, C.byref(gr), None)
    assert rc != 0 and b"dL_dshs must match shs" in lib.scgr_last_error()
    # null pointers
    assert lib.scgr_mark_visible(None, 3, None, None, None) != 0
    # the single-call forward needs its pinned status word; the same input validation applies
    rc = lib.scgr_forward(C.byref(v), C.byref(g), 256, 256, 256, 16, 256, 256, 256, 256, None, None)
    assert rc not in (0, _lib.NEED_CAPACITY) and b"pinned host status word" in lib.scgr_last_error()
    bad = ScgrGaussians(P=4, sh_coeffs=1, means3D=256, opacities=256, shs=256, colors_precomp=256,
                        scales=256, rotations=256, cov3D_precomp=None)
    rc = lib.scgr_forward(C.byref(v), C.byref(bad), 256, 256, 256, 16, 256, 256, 256, 256, 256, None)
    assert rc not in (0, _lib.NEED_CAPACITY) and b"exactly one of shs / colors_precomp" in lib.scgr_last_error()
    # photometric loss: shapes and pointers are checked before anything is launched
    assert lib.scgr_photometric_scratch_bytes(3, 1080, 1920) >= 3 * 4 * 3 * 1080 * 1920
    rc = lib.scgr_photometric_forward(256, 256, 3, 0, 16, 0.2, 256, 1, 256, None)
    assert rc != 0 and b"empty image" in lib.scgr_last_error()
    rc = lib.scgr_photometric_forward(256, None, 3, 16, 16, 0.2, 256, 1, 256, None)
    assert rc != 0 and b"null argument" in lib.scgr_last_error()
    rc = lib.scgr_photometric_backward(256, 256, 3, 16, 16, 0.2, None, None, 256, None)
    assert rc != 0 and b"null argument" in lib.scgr_last_error()


def test_argument_errors_of_the_model_pass_entry_points():
    """scgr_assemble_* / scgr_adam_step / scgr_densification_stats / scgr_gather_rows / scgr_copy_segments /
    scgr_knn3_mean_dist2: everything is validated before anything is launched, so this runs without a GPU."""
    from scgaussian_b200._lib import (ScgrActivated, ScgrAdamGroup, ScgrModel, ScgrModelSet, ScgrRowGather,
                                      ScgrSegmentCopy)
    lib = _lib.load()
    empty = ScgrModelSet(0, *[None] * 9)
    # a set needs exactly one position source, and every parameter array
    both = ScgrModelSet(4, 256, 256, 256, 256, 256, 256, 256, 256, 256)
    m = ScgrModel(15, (ScgrModelSet * 2)(both, empty))
    out = ScgrActivated(256, 256, 256, 256, 256)
    assert lib.scgr_assemble_forward(C.byref(m), C.byref(out), None) != 0
    assert b"either xyz or (rayo, rayd, zval)" in lib.scgr_last_error()
    no_rest = ScgrModelSet(4, 256, None, None, None, 256, 256, 256, 256, None)
    m = ScgrModel(15, (ScgrModelSet * 2)(empty, no_rest))
    assert lib.scgr_assemble_forward(C.byref(m), C.byref(out), None) != 0 and b"features_rest" in lib.scgr_last_error()
    unaligned = ScgrModelSet(4, 256, None, None, None, 256, 260, 256, 256, 256)
    m = ScgrModel(15, (ScgrModelSet * 2)(empty, unaligned))
    assert lib.scgr_assemble_forward(C.byref(m), C.byref(out), None) != 0 and b"16-byte aligned" in lib.scgr_last_error()
    m = ScgrModel(15, (ScgrModelSet * 2)(empty, empty))
    assert lib.scgr_assemble_forward(C.byref(m), C.byref(out), None) == 0          # an empty model is a no-op
    assert lib.scgr_assemble_forward(None, C.byref(out), None) != 0
    # Adam: table size, step counted from 1, null arrays
    g = (ScgrAdamGroup * 1)(ScgrAdamGroup(256, 256, 256, 256, 8, 0.1, 0))
    assert lib.scgr_adam_step(g, 1, 0.9, 0.999, 1e-15, None) != 0 and b"step counts from 1" in lib.scgr_last_error()
    g = (ScgrAdamGroup * 1)(ScgrAdamGroup(256, None, 256, 256, 8, 0.1, 1))
    assert lib.scgr_adam_step(g, 1, 0.9, 0.999, 1e-15, None) != 0 and b"null array" in lib.scgr_last_error()
    assert lib.scgr_adam_step(g, 17, 0.9, 0.999, 1e-15, None) != 0 and b"SCGR_ADAM_MAX_GROUPS" in lib.scgr_last_error()
    assert lib.scgr_adam_step(g, 1, 1.0, 0.999, 1e-15, None) != 0 and b"betas" in lib.scgr_last_error()
    assert lib.scgr_adam_step(None, 0, 0.9, 0.999, 1e-15, None) == 0
    g = (ScgrAdamGroup * 1)(ScgrAdamGroup(None, None, None, None, 0, 0.1, 1))
    assert lib.scgr_adam_step(g, 1, 0.9, 0.999, 1e-15, None) == 0                   # empty groups are skipped
    # statistics
    assert lib.scgr_densification_stats(256, None, None, 4, 256, 256, None, None) != 0
    assert b"update_filter or radii" in lib.scgr_last_error()
    assert lib.scgr_densification_stats(256, 256, None, 4, 256, None, None, None) != 0
    assert lib.scgr_densification_stats(None, None, None, 0, None, None, None, None) == 0
    # gather / copy
    a = (ScgrRowGather * 1)(ScgrRowGather(256, None, 3))
    assert lib.scgr_gather_rows(a, 1, 256, 5, None) != 0 and b"null array" in lib.scgr_last_error()
    assert lib.scgr_gather_rows(a, 49, 256, 5, None) != 0 and lib.scgr_gather_rows(a, 1, None, 5, None) != 0
    assert lib.scgr_gather_rows(a, 1, 256, 0, None) == 0
    c = (ScgrSegmentCopy * 1)(ScgrSegmentCopy(None, 256, 5))
    assert lib.scgr_copy_segments(c, 1, None) != 0 and b"null destination" in lib.scgr_last_error()
    c = (ScgrSegmentCopy * 1)(ScgrSegmentCopy(256, 256, -1))
    assert lib.scgr_copy_segments(c, 1, None) != 0 and lib.scgr_copy_segments(c, 49, None) != 0
    # distCUDA2
    assert lib.scgr_knn3_mean_dist2(None, 5, 256, None) != 0 and lib.scgr_knn3_mean_dist2(256, -1, 256, None) != 0
    assert lib.scgr_knn3_mean_dist2(None, 0, None, None) == 0
    assert lib.scgr_knn3_mean_dist2(256, (1 << 20) + 1, 256, None) != 0 and b"2^20" in lib.scgr_last_error()


def test_host_api_mirrors_reference_operator_surface():
    import diff_gaussian_rasterization as D
    from scgaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer, ScgrError
    assert D.GaussianRasterizer is GaussianRasterizer
    # the 12 keyword fields of reference gaussian_renderer/__init__.py:38-51
    s = GaussianRasterizationSettings(image_height=8, image_width=8, tanfovx=0.5, tanfovy=0.5,
                                      bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=torch.eye(4),
                                      projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3),
                                      prefiltered=False, debug=False)
    r = GaussianRasterizer(raster_settings=s)
    m3, m2, op = torch.zeros(5, 3), torch.zeros(5, 3), torch.ones(5, 1)
    sh, col = torch.zeros(5, 1, 3), torch.zeros(5, 3)
    sc, rot, cov = torch.ones(5, 3), torch.zeros(5, 4), torch.zeros(5, 6)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m3, means2D=m2, opacities=op, scales=sc, rotations=rot)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=m3, means2D=m2, opacities=op, shs=sh, colors_precomp=col, scales=sc, rotations=rot)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m3, means2D=m2, opacities=op, shs=sh)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc, rotations=rot, cov3D_precomp=cov)
    # no CPU path, ever: CPU tensors fail loudly instead of falling back
    with pytest.raises(ScgrError, match="CUDA"):
        r(means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc, rotations=rot)
    with pytest.raises(ScgrError, match="CUDA"):
        r.markVisible(m3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scgaussian_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# no oracle", ""), f"{f} mentions the oracle"
    txt = open(os.path.join(ROOT, "diff_gaussian_rasterization", "__init__.py")).read()
    assert "oracle" not in txt


def test_library_is_sm_100a_code_with_the_instructions_the_design_names():
    """DESIGN.md names the hardware paths of the kernels; this reads them back from the SASS of the shipped library: built for
    sm_100a only, TMA bulk copies + mbarrier waits (UBLKCP / SYNCS), cp.async staging (LDGSTS), multimem loads of the NVLS
    collective (LDGMC), fp32 reductions to global memory (REDG), programmatic dependent launch (ACQBULK = griddepcontrol.wait,
    PREEXIT = griddepcontrol.launch_dependents) -- and no kernel of a library (CUB / cuBLAS / Triton) inside it."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    for mnemonic, at_least in (("UBLKCP", 10), ("SYNCS", 10), ("LDGSTS", 10), ("LDGMC", 4), ("REDG.E.ADD.F32", 10),
                               ("ACQBULK", 20), ("PREEXIT", 20)):
        assert sass.count(mnemonic) >= at_least, (mnemonic, sass.count(mnemonic))
    kernels = re.findall(r"Function : (\S+)", sass)
    assert len(kernels) > 50
    foreign = [k for k in kernels if "scgr" not in k]
    assert not foreign, foreign[:5]
    # every kernel launched through chain() starts by waiting for its predecessor: one ACQBULK per instantiation
    chained = [k for k in kernels if re.search(r"onesweep_pass_kernel|emit_instances_kernel|render_forward_kernel|render_backward_kernel|"
                                               r"backward_prologue_kernel|preprocess_backward_kernel", k)]
    assert sass.count("ACQBULK") >= len(chained) > 20
