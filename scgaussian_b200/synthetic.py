"""Synthetic workload generator of SURVEY.md section 8d: the seed-0 scene, camera and upstream gradients that
bench.py, the tools and the parity tests all run on.  Host-side input construction only (CPU torch, no rasterizer
arithmetic): the camera matrices follow reference scene/cameras.py:54-63 and utils/graphics_utils.py:51-71
(pinned by tests/golden/reference_anchors.npz)."""
from __future__ import annotations

import math
from typing import Optional

import torch


def projection_matrix(znear, zfar, fovX, fovY, dtype=torch.float32):
    """Same entries as reference utils/graphics_utils.py:51-71 (checked by golden vectors)."""
    tY, tX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tY * znear, tX * znear
    bottom, left = -top, -right
    Pm = torch.zeros(4, 4, dtype=dtype)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_camera(W, H, fovx_deg=60.0, w2c: Optional[torch.Tensor] = None, znear=0.01, zfar=100.0,
                dtype=torch.float32):
    """Matrices built exactly as reference scene/cameras.py:54-63 (transposed / row-vector)."""
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * H / W
    fovx = 2 * math.atan(tanfovx)
    fovy = 2 * math.atan(tanfovy)
    if w2c is None:
        w2c = torch.eye(4, dtype=dtype)
    view = w2c.to(dtype).T.contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy, dtype).T
    full = view @ proj
    campos = torch.linalg.inv(view)[3, :3].contiguous()
    return dict(tanfovx=tanfovx, tanfovy=tanfovy, viewmatrix=view, projmatrix=full.contiguous(),
                campos=campos)


def yaw_w2c(deg: float) -> torch.Tensor:
    a = math.radians(deg)
    m = torch.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    return m


def synth_scene(P, W, H, sh_degree=3, scale_median=0.01, seed=0, fovx_deg=60.0,
                max_sh_degree=None, dtype=torch.float32):
    """SURVEY.md section 8d generator: draw order z, x, y, scale, quat, opacity, sh."""
    g = torch.Generator().manual_seed(seed)
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * H / W
    z = torch.rand(P, generator=g) * 8.0 + 2.0
    x = (torch.rand(P, generator=g) * 2 - 1) * tanfovx * z
    y = (torch.rand(P, generator=g) * 2 - 1) * tanfovy * z
    means3D = torch.stack([x, y, z], -1)
    # exp() on one thread: with several, the chunking (hence which elements take the vector and
    # which the scalar code path, 1 ulp apart) varies from process to process, and so would R
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        scales = torch.exp(torch.randn(P, 3, generator=g) * 0.5 + math.log(scale_median))
    finally:
        torch.set_num_threads(nt)
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.rand(P, 1, generator=g) * 0.9 + 0.05
    M = ((max_sh_degree if max_sh_degree is not None else sh_degree) + 1) ** 2
    shs = torch.randn(P, M, 3, generator=g) * 0.1
    shs[:, 0] *= 5.0     # DC ~ N(0, 0.5^2), rest ~ N(0, 0.1^2)
    return dict(means3D=means3D.to(dtype), scales=scales.to(dtype), rotations=rotations.to(dtype),
                opacities=opacities.to(dtype), shs=shs.to(dtype))


def synth_upstream_grads(W, H, seed=1, dtype=torch.float32):
    """dL/dcolor, dL/ddepth, dL/dalpha = N(0,1)/N."""
    g = torch.Generator().manual_seed(seed)
    N = W * H
    return (torch.randn(3, H, W, generator=g).to(dtype) / N,
            torch.randn(1, H, W, generator=g).to(dtype) / N,
            torch.randn(1, H, W, generator=g).to(dtype) / N)
