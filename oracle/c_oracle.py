"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/scg_oracle.c (see that file's header:
parity unpinned; restates SURVEY.md Appendix A).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "scg_oracle.c")
    libs = [os.path.join(_HERE, f"liboracle_{p}.so") for p in ("f32", "f64")]
    if force or any((not os.path.exists(l)) or os.path.getmtime(l) < os.path.getmtime(src) for l in libs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])


class _Inputs32(C.Structure):
    pass


def _make_struct(real):
    ptr = C.POINTER(real)

    class Inputs(C.Structure):
        _fields_ = [("P", C.c_int), ("M", C.c_int), ("D", C.c_int), ("W", C.c_int), ("H", C.c_int),
                    ("tanfovx", real), ("tanfovy", real), ("scale_modifier", real),
                    ("bg", ptr), ("viewmatrix", ptr), ("projmatrix", ptr), ("campos", ptr),
                    ("means3D", ptr), ("opacities", ptr), ("shs", ptr), ("colors_precomp", ptr),
                    ("scales", ptr), ("rotations", ptr), ("cov3D_precomp", ptr)]
    return Inputs


class COracle:
    """One forward (+ optional backward) of the scalar CPU rasterizer.  Arrays are numpy."""

    def __init__(self, precision: str = "f32"):
        build()
        self.np_real = np.float32 if precision == "f32" else np.float64
        self.c_real = C.c_float if precision == "f32" else C.c_double
        self.lib = C.CDLL(os.path.join(_HERE, f"liboracle_{precision}.so"))
        assert self.lib.scgo_real_size() == np.dtype(self.np_real).itemsize
        self.Inputs = _make_struct(self.c_real)
        self.lib.scgo_forward.restype = C.c_void_p
        self.lib.scgo_forward_ex.restype = C.c_void_p
        self.lib.scgo_num_rendered.restype = C.c_int64
        for f in ("scgo_point_list", "scgo_ranges", "scgo_means2D", "scgo_conic", "scgo_rgb",
                  "scgo_depths", "scgo_tiles_touched", "scgo_n_contrib", "scgo_geom_margin"):
            getattr(self.lib, f).restype = C.c_void_p
        self._state = None
        self._keep = None

    def set_threads(self, n: int) -> int:
        """OpenMP threads of the following calls (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
        return int(self.lib.scgo_set_threads(int(n)))

    def _arr(self, a, shape=None):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, dtype=self.np_real))
        return a

    def _p(self, a):
        return a.ctypes.data_as(C.POINTER(self.c_real)) if a is not None else None

    def forward(self, *, means3D, opacities, W, H, tanfovx, tanfovy, bg, viewmatrix, projmatrix,
                campos, sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, scale_modifier=1.0, override=None):
        """override: dict(xy [P,2], conic [P,3], rgb [P,3], depth [P], radii [P] int32, clamped [P,3] uint8) -- per-
        Gaussian 2D state used instead of the oracle's own preprocess (staged parity, scg_oracle.c: scgo_override)."""
        self.free()
        a = {k: self._arr(v) for k, v in dict(
            bg=bg, viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos, means3D=means3D,
            opacities=opacities, shs=shs, colors_precomp=colors_precomp, scales=scales,
            rotations=rotations, cov3D_precomp=cov3D_precomp).items()}
        P = a["means3D"].shape[0]
        M = a["shs"].shape[1] if a["shs"] is not None else 0
        inp = self.Inputs(P, M, int(sh_degree), int(W), int(H), tanfovx, tanfovy, scale_modifier,
                          *[self._p(a[k]) for k in ("bg", "viewmatrix", "projmatrix", "campos",
                                                    "means3D", "opacities", "shs", "colors_precomp",
                                                    "scales", "rotations", "cov3D_precomp")])
        color = np.zeros((3, H, W), self.np_real)
        depth = np.zeros((1, H, W), self.np_real)
        alpha = np.zeros((1, H, W), self.np_real)
        radii = np.zeros(P, np.int32)
        ovp = None
        ov_keep = None
        if override is not None:
            rp = C.POINTER(self.c_real)

            class Override(C.Structure):
                _fields_ = [("xy", rp), ("conic", rp), ("rgb", rp), ("depth", rp), ("radii", C.POINTER(C.c_int)),
                            ("clamped", C.POINTER(C.c_ubyte))]
            ov_keep = {k: self._arr(override[k]) for k in ("xy", "conic", "rgb", "depth")}
            ov_keep["radii"] = np.ascontiguousarray(override["radii"], dtype=np.int32)
            ov_keep["clamped"] = np.ascontiguousarray(override["clamped"], dtype=np.uint8)
            assert ov_keep["xy"].shape == (P, 2) and ov_keep["conic"].shape == (P, 3) and ov_keep["radii"].shape == (P,)
            ov = Override(*[self._p(ov_keep[k]) for k in ("xy", "conic", "rgb", "depth")],
                          ov_keep["radii"].ctypes.data_as(C.POINTER(C.c_int)),
                          ov_keep["clamped"].ctypes.data_as(C.POINTER(C.c_ubyte)))
            ovp = C.byref(ov)
        self._keep = (a, inp, ov_keep)
        self._state = C.c_void_p(self.lib.scgo_forward_ex(
            C.byref(inp), ovp, self._p(color), self._p(depth), self._p(alpha),
            radii.ctypes.data_as(C.POINTER(C.c_int))))
        self._dims = (P, M, W, H)
        return color, radii, depth, alpha

    def backward(self, dL_dcolor, dL_ddepth, dL_dalpha):
        P, M, W, H = self._dims
        gC, gD, gA = self._arr(dL_dcolor), self._arr(dL_ddepth), self._arr(dL_dalpha)
        z = lambda *s: np.zeros(s, self.np_real)
        out = dict(means3D=z(P, 3), means2D=z(P, 3), shs=z(P, max(M, 1), 3), colors_precomp=z(P, 3),
                   opacities=z(P, 1), scales=z(P, 3), rotations=z(P, 4), cov3D_precomp=z(P, 6))
        self.lib.scgo_backward(self._state, self._p(gC), self._p(gD), self._p(gA),
                               *[self._p(out[k]) for k in ("means3D", "means2D", "shs", "colors_precomp",
                                                           "opacities", "scales", "rotations",
                                                           "cov3D_precomp")])
        if M == 0:
            out["shs"] = z(P, 0, 3)
        return out

    # --- intermediate state, for stage-level checks of the CUDA path -------------------
    def _view(self, fn, dtype, n):
        addr = getattr(self.lib, fn)(self._state)
        if n == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype).copy()

    @property
    def num_rendered(self):
        return int(self.lib.scgo_num_rendered(self._state))

    def state(self):
        P, M, W, H = self._dims
        Tn = ((W + 15) // 16) * ((H + 15) // 16)
        R = self.num_rendered
        return dict(point_list=self._view("scgo_point_list", np.uint32, R),
                    ranges=self._view("scgo_ranges", np.int64, 2 * Tn).reshape(Tn, 2),
                    means2D=self._view("scgo_means2D", self.np_real, 2 * P).reshape(P, 2),
                    conic=self._view("scgo_conic", self.np_real, 3 * P).reshape(P, 3),
                    rgb=self._view("scgo_rgb", self.np_real, 3 * P).reshape(P, 3),
                    depths=self._view("scgo_depths", self.np_real, P),
                    tiles_touched=self._view("scgo_tiles_touched", np.int32, P),
                    n_contrib=self._view("scgo_n_contrib", np.int32, W * H).reshape(H, W))

    def margins(self, base_err=2e-6, conic_err=2e-6, pos_ulps=4.0, geom_err=None):
        """Which outputs a flipped discrete decision could touch (scg_oracle.c: scgo_margins, error model there).
        pos_ulps: uncertainty of a projected mean in units of 2^-23 x max(W, H) pixels.  Returns
        pix_margin [H,W] (margin / uncertainty, < 1 <=> flip-prone), pix_flag [H,W] bool, gauss_margin [P],
        gauss_flag [P] bool (flip-affected), gauss_own [P] bool (the uncertain decision is the Gaussian's own),
        geom_margin [P] (pixels)."""
        P, M, W, H = self._dims
        pos_err = pos_ulps * 2.0 ** -23 * max(W, H)
        geom_err = 4.0 * pos_err if geom_err is None else geom_err
        pm = np.zeros((H, W), self.np_real)
        pf = np.zeros((H, W), np.uint8)
        gm = np.zeros(max(P, 1), self.np_real)
        gf = np.zeros(max(P, 1), np.uint8)
        self.lib.scgo_margins(self._state, C.c_double(base_err), C.c_double(conic_err), C.c_double(pos_err),
                              C.c_double(geom_err), self._p(pm), pf.ctypes.data_as(C.POINTER(C.c_ubyte)), self._p(gm),
                              gf.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return dict(pix_margin=pm, pix_flag=pf.astype(bool), gauss_margin=gm[:P], gauss_flag=gf[:P] > 0,
                    gauss_own=gf[:P] == 2, geom_margin=self._view("scgo_geom_margin", self.np_real, P),
                    model=dict(base_err=base_err, conic_err=conic_err, pos_err_px=pos_err, geom_err_px=geom_err))

    def free(self):
        if self._state is not None:
            self.lib.scgo_free(self._state)
            self._state = None
            self._keep = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
