"""BASELINE / TEST INFRASTRUCTURE -- ctypes front-end of baseline/standin/standin.cu: the deliberately naive CUDA
restatement of the published rasterizer algorithm (see that file's header).  NOT the reference and NOT the product;
nothing under scgaussian_b200/ imports it.  bench.py times it as `gpu_standin_baseline`; tests/test_standin.py
checks it against the CPU oracle."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "standin.cu")
LIB = os.path.join(_HERE, "libstandin.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                               "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC], env=env)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        _lib.standin_forward.restype = C.c_int
        _lib.standin_forward.argtypes = [i, i, i, i, i, f, f, f] + [vp] * 13 + [C.POINTER(C.c_longlong), vp]
        _lib.standin_backward.restype = C.c_int
        _lib.standin_backward.argtypes = [vp] * 10
    return _lib


class Standin:
    """One scene resident on `dev`; forward() / backward() mirror the product's raw stage calls."""

    def __init__(self, scene, W, H, sh_degree, dev):
        self.lib = load()
        self.dev, self.W, self.H, self.D = dev, W, H, sh_degree
        self.t = {k: scene[k].to(dev).float().contiguous() for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        self.P = int(self.t["means3D"].shape[0])
        self.M = int(self.t["shs"].shape[1])
        P, M = self.P, self.M
        z = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)     # noqa: E731
        self.color, self.depth, self.alpha = z(3, H, W), z(1, H, W), z(1, H, W)
        self.radii = torch.empty(P, dtype=torch.int32, device=dev)
        self.g = dict(means3D=z(P, 3), means2D=z(P, 3), shs=z(P, M, 3), opacities=z(P, 1), scales=z(P, 3), rotations=z(P, 4))
        self.bg = torch.zeros(3, device=dev)
        self.num_rendered = 0

    def forward(self, cam, bg=None, scale_modifier=1.0):
        keep = [cam["viewmatrix"].to(self.dev).float().contiguous(), cam["projmatrix"].to(self.dev).float().contiguous(),
                cam["campos"].to(self.dev).float().contiguous(), (self.bg if bg is None else bg.to(self.dev).float().contiguous())]
        self._keep = keep
        R = C.c_longlong(0)
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        t = self.t
        rc = self.lib.standin_forward(self.P, self.M, self.D, self.W, self.H, cam["tanfovx"], cam["tanfovy"], scale_modifier,
                                      keep[3].data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(),
                                      t["means3D"].data_ptr(), t["opacities"].data_ptr(), t["shs"].data_ptr(),
                                      t["scales"].data_ptr(), t["rotations"].data_ptr(), self.color.data_ptr(),
                                      self.depth.data_ptr(), self.alpha.data_ptr(), self.radii.data_ptr(), C.byref(R), st)
        if rc != 0:
            raise RuntimeError(f"standin_forward failed ({rc})")
        self.num_rendered = int(R.value)
        return self.color, self.radii, self.depth, self.alpha

    def backward(self, gC, gD, gA):
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        g = self.g
        rc = self.lib.standin_backward(gC.data_ptr(), gD.data_ptr(), gA.data_ptr(), g["means3D"].data_ptr(), g["means2D"].data_ptr(),
                                       g["shs"].data_ptr(), g["opacities"].data_ptr(), g["scales"].data_ptr(),
                                       g["rotations"].data_ptr(), st)
        if rc != 0:
            raise RuntimeError(f"standin_backward failed ({rc})")
        return g


def measure(scene_cpu, cams_cpu, grads_cpu, W, H, sh_degree, dev, steps=8, warm=2):
    """views/s forward + backward of the stand-in on the bench's own inputs, a different camera every step."""
    sb = Standin(scene_cpu, W, H, sh_degree, dev)
    gC, gD, gA = [g.to(dev).float().contiguous() for g in grads_cpu]
    Rs = []
    for i in range(warm):
        sb.forward(cams_cpu[i % len(cams_cpu)])
        sb.backward(gC, gD, gA)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        sb.forward(cams_cpu[(warm + i) % len(cams_cpu)])
        Rs.append(sb.num_rendered)
        sb.backward(gC, gD, gA)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    load().standin_release()
    return {"what": "naive CUDA restatement of the published algorithm (baseline/standin/standin.cu): thread-per-Gaussian "
                    "preprocess, cub::DeviceScan + blocking D2H of R, cub::DeviceRadixSort on 64-bit tile|depth keys, one "
                    "256-thread CTA per tile, 10 global atomics per contributing pair; checked against the CPU oracle in "
                    "tests/test_standin.py; NOT the reference (unobtainable offline)",
            "ms_per_step": ms, "views_s": 1000.0 / ms, "steps": steps, "num_rendered_R_mean": sum(Rs) / len(Rs)}
