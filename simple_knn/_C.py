"""`simple_knn._C` as the reference imports it (scene/gaussian_model.py:20): distCUDA2(points) -> [N]
mean squared distance of every point to its 3 nearest neighbours (used once, at :444, to size the initial
Gaussians).  Computes in libscgr.so through the C ABI; no CPU path."""
import ctypes as C

import torch

from scgaussian_b200 import _lib
from scgaussian_b200._lib import ScgrError, check


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if points.device.type != "cuda":
        raise ScgrError("distCUDA2 runs on CUDA tensors only (no CPU path exists)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise ScgrError("distCUDA2 expects points of shape [N, 3]")
    lib = _lib.load()
    pts = points.detach().float().contiguous()
    n = int(pts.shape[0])
    out = torch.empty(n, dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        stream = C.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream)
        check(lib.scgr_knn3_mean_dist2(pts.data_ptr(), n, out.data_ptr(), stream))
    return out
