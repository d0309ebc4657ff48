"""ctypes binding of libscgr.so (include/scgr.h).  This is the ONLY route from Python to the
rasterizer: there is no eager / CPU fallback -- if the library is missing or fails to load, every
operator call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCGR_LIB selects another build of the same library (compile-time A/B variants, tools/ab.py)
LIB_PATH = os.environ.get("SCGR_LIB") or os.path.join(_HERE, "libscgr.so")


class ScgrView(C.Structure):
    _fields_ = [("image_height", C.c_int32), ("image_width", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("bg", C.c_void_p), ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
                ("sh_degree", C.c_int32), ("campos", C.c_void_p),
                ("prefiltered", C.c_int32), ("debug", C.c_int32)]


class ScgrGaussians(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_coeffs", C.c_int32),
                ("means3D", C.c_void_p), ("opacities", C.c_void_p), ("shs", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("cov3D_precomp", C.c_void_p), ("sh_dc", C.c_void_p * 2), ("sh_rest", C.c_void_p * 2), ("sh_n0", C.c_int32)]


class ScgrGrads(C.Structure):
    _fields_ = [("dL_dmeans3D", C.c_void_p), ("dL_dmeans2D", C.c_void_p), ("dL_dshs", C.c_void_p),
                ("dL_dcolors_precomp", C.c_void_p), ("dL_dopacities", C.c_void_p),
                ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
                ("dL_dcov3D_precomp", C.c_void_p), ("densification_stats", C.c_void_p), ("radii", C.c_void_p),
                ("accumulate", C.c_int32), ("live_count", C.c_void_p), ("dL_dsh_dc", C.c_void_p * 2),
                ("dL_dsh_rest", C.c_void_p * 2)]


class ScgrNvlsFused(C.Structure):
    _fields_ = [("multicast_ptr", C.c_void_p), ("dense_floats", C.c_size_t), ("multicast_rows", C.c_void_p),
                ("live_count", C.c_void_p), ("n_rows", C.c_int64), ("row_floats", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("flags", C.c_void_p * 8), ("sync_local", C.c_void_p), ("epoch", C.c_uint32),
                ("peer_ptrs", C.c_void_p * 8)]


class ScgrDebugViews(C.Structure):
    _fields_ = [("record", C.c_void_p), ("tiles_touched", C.c_void_p), ("depth_order", C.c_void_p),
                ("point_list", C.c_void_p), ("ranges", C.c_void_p), ("n_contrib", C.c_void_p),
                ("final_T", C.c_void_p), ("num_rendered", C.c_void_p)]


class ScgrMatchPair(C.Structure):
    _fields_ = [("n", C.c_int32), ("uv0", C.c_void_p), ("rays_o", C.c_void_p), ("rays_d", C.c_void_p),
                ("cam_rays_d", C.c_void_p), ("uv1", C.c_void_p), ("valid", C.c_void_p), ("w2c1", C.c_float * 12),
                ("intr1", C.c_float * 9)]


MATCH_MAX_PAIRS = 8    # SCGR_MATCH_MAX_PAIRS (include/scgr.h)


class ScgrModelSet(C.Structure):
    _fields_ = [("n", C.c_int32), ("xyz", C.c_void_p), ("rayo", C.c_void_p), ("rayd", C.c_void_p),
                ("zval", C.c_void_p), ("scaling", C.c_void_p), ("rotation", C.c_void_p), ("opacity", C.c_void_p),
                ("features_dc", C.c_void_p), ("features_rest", C.c_void_p)]


class ScgrModel(C.Structure):
    _fields_ = [("sh_rest", C.c_int32), ("set", ScgrModelSet * 2)]


class ScgrActivated(C.Structure):
    _fields_ = [("means3D", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("opacities", C.c_void_p), ("shs", C.c_void_p)]


class ScgrActivatedGrads(C.Structure):
    _fields_ = [("dL_dmeans3D", C.c_void_p), ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
                ("dL_dopacities", C.c_void_p), ("dL_dshs", C.c_void_p)]


class ScgrModelSetGrads(C.Structure):
    _fields_ = [("dL_dxyz", C.c_void_p), ("dL_dzval", C.c_void_p), ("dL_dscaling", C.c_void_p),
                ("dL_drotation", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dfeatures_dc", C.c_void_p),
                ("dL_dfeatures_rest", C.c_void_p)]


class ScgrModelGrads(C.Structure):
    _fields_ = [("set", ScgrModelSetGrads * 2)]


class ScgrAdamGroup(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_double), ("step", C.c_int32)]


class ScgrRowGather(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("row_floats", C.c_int32)]


class ScgrSegmentCopy(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("n_floats", C.c_int64)]


COPY_MAX_SEGMENTS = 48   # SCGR_COPY_MAX_SEGMENTS (include/scgr.h)
GATHER_MAX_ARRAYS = 48   # SCGR_GATHER_MAX_ARRAYS (include/scgr.h)
ADAM_MAX_GROUPS = 16   # SCGR_ADAM_MAX_GROUPS (include/scgr.h)

# every symbol include/scgr.h declares: (restype, argtypes)
SYMBOLS = {
    "scgr_version": (C.c_int, []),
    "scgr_last_error": (C.c_char_p, []),
    "scgr_geometry_bytes": (C.c_size_t, [C.c_int32]),
    "scgr_binning_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "scgr_image_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "scgr_forward_geometry": (C.c_int, [C.POINTER(ScgrView), C.POINTER(ScgrGaussians), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_forward_render": (C.c_int, [C.POINTER(ScgrView), C.POINTER(ScgrGaussians), C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_forward": (C.c_int, [C.POINTER(ScgrView), C.POINTER(ScgrGaussians), C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "scgr_backward": (C.c_int, [C.POINTER(ScgrView), C.POINTER(ScgrGaussians), C.c_void_p, C.c_void_p,
                                C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(ScgrGrads), C.c_void_p]),
    "scgr_photometric_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "scgr_photometric_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "scgr_photometric_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_match_loss_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.POINTER(ScgrMatchPair),
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_match_loss_backward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.POINTER(ScgrMatchPair),
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_bg_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "scgr_masked_mean_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "scgr_masked_mean_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_masked_mean_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_nvls_allreduce": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p]),
    "scgr_nvls_allreduce_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "scgr_nvls_allreduce_fused": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scgr_knn3_mean_dist2": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "scgr_assemble_forward": (C.c_int, [C.POINTER(ScgrModel), C.POINTER(ScgrActivated), C.c_void_p]),
    "scgr_assemble_backward": (C.c_int, [C.POINTER(ScgrModel), C.POINTER(ScgrActivatedGrads),
                                         C.POINTER(ScgrModelGrads), C.c_void_p]),
    "scgr_densification_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "scgr_gather_rows": (C.c_int, [C.POINTER(ScgrRowGather), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "scgr_copy_segments": (C.c_int, [C.POINTER(ScgrSegmentCopy), C.c_int32, C.c_void_p]),
    "scgr_adam_step": (C.c_int, [C.POINTER(ScgrAdamGroup), C.c_int32, C.c_double, C.c_double, C.c_double,
                                 C.c_void_p]),
    "scgr_mark_visible": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scgr_kernel_launch_count": (C.c_longlong, []),
    "scgr_profile_enable": (C.c_int, [C.c_int]),
    "scgr_profile_fetch": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]),
    "scgr_debug_views": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(ScgrDebugViews)]),
}

NEED_CAPACITY = 3   # SCGR_NEED_CAPACITY (include/scgr.h)

_lib = None


class ScgrError(RuntimeError):
    pass


def load():
    """Loads libscgr.so (once).  Raises ScgrError -- never falls back -- when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScgrError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python -m scgaussian_b200.build`). There is no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise ScgrError(load().scgr_last_error().decode("utf-8", "replace"))
