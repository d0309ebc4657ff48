"""View-sharded data parallelism for the rasterizer path: parameters replicated,
rank k renders view k of the batch, gradients are summed with ONE all-reduce per step.

The reference is single-GPU (reference utils/general_utils.py:139) and trains one view per step
(reference train.py:135); summing per-view gradients over an N-view batch is mathematically the
gradient accumulation of N of its steps.  The path has no other exchange, so this is the only
collective: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch on the GPU box,
gloo in the CPU tests) over ONE flat fp32 buffer that the backward kernels write into directly
(scgr_backward fills every gradient tensor in full, so the buffer needs no zeroing).

Buffer layout (struct-of-arrays, each block a contiguous [P, k] tensor the C ABI can write):
    means3D 3 | shs 3M (or colors 3) | opacities 1 | scales 3 | rotations 4 (or cov3D 6) | stats 2
`stats` carries the densification statistics the reference accumulates per view at
reference scene/gaussian_model.py:932-934: ||dL/dmean2D[:, :2]|| * visible and visible, so that
the single SUM all-reduce yields exactly what N sequential add_densification_stats calls would.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, P: int, sh_coeffs: int = 16, use_sh: bool = True, use_cov: bool = False,
                 device="cuda", with_stats: bool = True):
        self.P = P
        fields = [("means3D", (P, 3))]
        fields.append(("shs", (P, sh_coeffs, 3)) if use_sh else ("colors_precomp", (P, 3)))
        fields.append(("opacities", (P, 1)))
        if use_cov:
            fields.append(("cov3D_precomp", (P, 6)))
        else:
            fields += [("scales", (P, 3)), ("rotations", (P, 4))]
        if with_stats:
            fields.append(("stats", (P, 2)))
        self.fields = fields
        sizes = []
        for _, shp in fields:
            n = 1
            for d in shp:
                n *= d
            sizes.append((n + 3) // 4 * 4)          # keep every block 16-byte aligned
        self.flat = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        o = 0
        for (name, shp), n in zip(fields, sizes):
            numel = 1
            for d in shp:
                numel *= d
            self.views[name] = self.flat[o:o + numel].view(*shp)
            o += n
        # means2D is returned by the op but is not a parameter gradient: keep it outside the flat buffer
        self.means2D = torch.empty(P, 3, dtype=torch.float32, device=device)

    def out_dict(self) -> Dict[str, torch.Tensor]:
        d = {k: v for k, v in self.views.items() if k != "stats"}
        d["means2D"] = self.means2D
        return d

    def fill_stats(self, radii: torch.Tensor) -> None:
        """reference scene/gaussian_model.py:932-934 for this rank's view."""
        if "stats" not in self.views:
            return
        vis = (radii > 0).to(torch.float32)
        st = self.views["stats"]
        st[:, 0] = torch.linalg.vector_norm(self.means2D[:, :2], dim=-1) * vis
        st[:, 1] = vis

    def all_reduce(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """THE collective of the path: one SUM all-reduce over the flat buffer."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    def nbytes(self) -> int:
        return self.flat.numel() * 4


def shard_views(n_views: int, rank: int, world_size: int):
    """Views of a batch handled by `rank`: round-robin, one view per GPU when n_views == world_size."""
    return list(range(rank, n_views, world_size))
