"""View-sharded data parallelism for the rasterizer path: parameters replicated,
rank k renders view k of the batch, gradients are summed with ONE all-reduce per step.

The reference is single-GPU (reference utils/general_utils.py:139) and trains one view per step
(reference train.py:135); summing per-view gradients over an N-view batch is mathematically the
gradient accumulation of N of its steps.  The path has no other exchange, so this is the only
collective: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch on the GPU box,
gloo in the CPU tests) over ONE flat fp32 buffer that the backward kernels write into directly
(scgr_backward fills every gradient tensor in full, so the buffer needs no zeroing).

On a box whose GPUs share an NVSwitch the flat buffer is allocated in symmetric memory with a multicast
mapping (torch.distributed._symmetric_memory: plumbing) and the sum is done by libscgr's own two-shot
NVLS kernel (scgr_nvls_allreduce, csrc/collective.cu: in-switch multimem.ld_reduce of this rank's 1/N of
the buffer, multimem.st broadcast), bracketed by the symmetric-memory barriers.  Anything that prevents
that (CPU tensors, no multicast support, fewer than 4 ranks -- where NCCL's direct P2P path is faster --,
SCGR_ALLREDUCE=nccl) leaves the collective to torch.distributed.all_reduce on the same buffer.

Buffer layout (struct-of-arrays, each block a contiguous [P, k] tensor the C ABI can write):
    means3D 3 | colors 3 (no-SH path) | opacities 1 | scales 3 | rotations 4 (or cov3D 6) | stats 2 | live 1 | shs 3M
`stats` carries the densification statistics the reference accumulates per view at
reference scene/gaussian_model.py:932-934: ||dL/dmean2D[:, :2]|| * visible and visible, so that
the single SUM all-reduce yields exactly what N sequential add_densification_stats calls would.
`live` counts, per Gaussian, the views that gave it any gradient (ScgrGrads.live_count).  In one view most Gaussians
receive none (hidden behind saturated pixels, or out of view): their gradient rows are exact zeros on every rank.  The
NVLS path therefore reduces the small blocks + `live` densely, and then the 192-byte dL/dSH rows -- 79 % of the bytes
-- only for the Gaussians some rank marked live (scgr_nvls_allreduce_rows): same result as the dense sum, bit for
bit, a fraction of the traffic.  The NCCL path reduces the whole buffer in one dist.all_reduce.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist


class FlatGradBuffer:
    FLAG_WORDS = 64          # >= world; keeps the allocation a multiple of 256 bytes
    P2P_MAX_WORLD = 4        # largest world that hands the kernel the peer mappings (it goes peer to peer for both shots at 2
                             # ranks, for the dense one at 4: measured, profiles/r02_scaling.md)

    def __init__(self, P: int, sh_coeffs: int = 16, use_sh: bool = True, use_cov: bool = False,
                 device="cuda", with_stats: bool = True, symmetric: Optional[bool] = None):
        self.P = P
        self._symm = None            # symmetric-memory handle when the NVLS path is active
        fields = [("means3D", (P, 3))]
        if not use_sh:
            fields.append(("colors_precomp", (P, 3)))
        fields.append(("opacities", (P, 1)))
        if use_cov:
            fields.append(("cov3D_precomp", (P, 6)))
        else:
            fields += [("scales", (P, 3)), ("rotations", (P, 4))]
        if with_stats:
            fields.append(("stats", (P, 2)))
        fields.append(("live", (P,)))
        if use_sh:
            fields.append(("shs", (P, sh_coeffs, 3)))       # last: the block the row-sparse shot covers
        self.fields = fields
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        sizes = []
        for name, shp in fields:
            n = 1
            for d in shp:
                n *= d
            sizes.append((n + 3) // 4 * 4)          # keep every block 16-byte aligned
        # the dense part (everything before shs) is a whole number of float4 per rank
        dense = sum(sizes[:-1]) if use_sh else sum(sizes)
        pad = (-dense) % (4 * world)
        if use_sh:
            sizes[-2] += pad
        else:
            sizes[-1] += pad
        total = sum(sizes)
        total = (total + 4 * world - 1) // (4 * world) * (4 * world)      # every rank reduces an equal float4 shard
        self.dense_floats = dense + pad
        self.row_floats = 3 * sh_coeffs if use_sh and (3 * sh_coeffs) % 4 == 0 else 0      # 0: no row-sparse shot
        # behind the buffer, in the same symmetric allocation: the flag words of the fused collective's in-kernel barriers
        storage = self._allocate(total + self.FLAG_WORDS, torch.device(device), world, symmetric)
        self.flat = storage[:total]
        self._flag_offset = total
        self._epoch = 1
        self._sync = None
        if self._symm is not None:
            storage[total:].zero_()
            self._sync = torch.zeros(4, dtype=torch.int32, device=storage.device)
            torch.cuda.current_stream(storage.device).synchronize()
            self._symm.barrier(channel=0)        # every rank's flag words are zero before anyone's first collective
        self._storage = storage
        self.views: Dict[str, torch.Tensor] = {}
        o = 0
        for (name, shp), n in zip(fields, sizes):
            numel = 1
            for d in shp:
                numel *= d
            self.views[name] = self.flat[o:o + numel].view(*shp)
            if name == "shs":
                self.rows_offset = o
            o += n
            self.payload_floats = o - n + numel      # end of the last field: what follows is padding up to 4 * world floats
        # means2D is returned by the op but is not a parameter gradient: keep it outside the flat buffer
        self.means2D = torch.empty(P, 3, dtype=torch.float32, device=device)

    def out_dict(self) -> Dict[str, torch.Tensor]:
        """What rasterize_backward_raw(out=...) writes into: the parameter gradients AND the per-view densification
        statistics (scgr_backward produces both: ScgrGrads.densification_stats), all inside the flat buffer."""
        d = dict(self.views)
        d["means2D"] = self.means2D
        return d

    def fresh_out_dict(self) -> Dict[str, torch.Tensor]:
        """out_dict() made of NEW view tensors of the same memory: what the autograd node hands back as gradients, so
        that autograd can adopt them as `.grad` without a copy (it clones a gradient somebody else still references)."""
        d = {name: v.view(v.shape) for name, v in self.views.items()}
        d["means2D"] = self.means2D.view(self.means2D.shape)
        return d

    def capture(self):
        """Context manager: while active, the PUBLIC operator's backward (GaussianRasterizer / autograd) writes the
        parameter gradients, the densification statistics and the live counts of the view straight into this buffer --
        `loss.backward()` leaves it ready for `all_reduce()`, with no packing copies; afterwards the summed gradients are
        in `self.views[...]` (a leaf's `.grad` aliases its view whenever autograd could adopt the tensor)."""
        import contextlib
        from . import rasterizer as R

        @contextlib.contextmanager
        def cm():
            prev = getattr(R._grad_sink, "fn", None)
            R.set_gradient_sink(self.fresh_out_dict)
            try:
                yield self
            finally:
                R.set_gradient_sink(prev)
        return cm()

    def fill_stats(self, radii: torch.Tensor) -> None:
        """reference scene/gaussian_model.py:932-934 for this rank's view, from torch ops -- for callers whose backward
        did not go through out_dict() (e.g. autograd through the public operator); scgr_backward writes the same two
        columns itself when given the `stats` view."""
        if "stats" not in self.views:
            return
        vis = (radii > 0).to(torch.float32)
        st = self.views["stats"]
        st[:, 0] = torch.linalg.vector_norm(self.means2D[:, :2], dim=-1) * vis
        st[:, 1] = vis

    def fill_live(self, dL_dopacities: torch.Tensor, dL_dmeans3D: Optional[torch.Tensor] = None) -> None:
        """The live counts for callers whose backward did not go through out_dict() (autograd through the public
        operator, which hands back plain gradient tensors): a Gaussian counts as live when its dL/dopacity or any
        component of its dL/dmean3D is non-zero.  (scgr_backward itself marks a Gaussian live when ANY of its ten
        screen-space sums is non-zero; the two criteria differ only if those sums cancel to exactly 0.0 in both the
        opacity and all three position components while another one does not.)"""
        live = dL_dopacities.reshape(-1) != 0
        if dL_dmeans3D is not None:
            live = live | (dL_dmeans3D != 0).any(dim=1)
        self.views["live"].copy_(live.to(torch.float32))

    def _allocate(self, total: int, device: torch.device, world: int, symmetric: Optional[bool]) -> torch.Tensor:
        # measured on 8xB200, dense 244 MB: NVLS kernel 609 us vs NCCL 661 us at 8 ranks, 645 us vs 481 us at 2 ranks;
        # with the row-sparse second shot the switch moves a fraction of that: it is used whenever it is available
        mode = os.environ.get("SCGR_ALLREDUCE", "auto")
        want = symmetric if symmetric is not None else (mode == "nvls" or (mode == "auto" and world >= 2))
        if want and world > 1 and device.type == "cuda" and dist.get_backend() == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                flat = symm_mem.empty(total, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
                if int(hdl.multicast_ptr) != 0:
                    self._symm = hdl
                    return flat
            except Exception as e:          # no symmetric memory / no multicast on this box: NCCL does the sum
                if symmetric:
                    raise
                self._symm_error = repr(e)
        return torch.empty(total, dtype=torch.float32, device=device)

    @property
    def collective(self) -> str:
        if self._symm is None:
            return "nccl all_reduce"
        w = int(self._symm.world_size)
        how = ", one launch with in-kernel barriers" if self._fused(w) else ""
        if self._fused(w) and self._p2p(w):
            how += ", peer-to-peer loads / stores" + ("" if w <= 2 or os.environ.get("SCGR_NVLS_P2P") == "1" else " for the dense shot")
        return ("nvls two-shot kernels (libscgr): dense small blocks + row-sparse dL/dSH" if self._sparse() else
                "nvls two-shot kernel (libscgr), dense") + how

    def _fused(self, world: int) -> bool:
        return world <= 8 and os.environ.get("SCGR_ALLREDUCE_FUSED", "1") != "0"

    def _p2p(self, world: int) -> bool:
        """Plain peer loads / stores instead of the multicast instructions: measured faster with few ranks (2: 0.19 ->
        ~0.1 ms for the small blocks).  SCGR_NVLS_P2P=auto|0|1."""
        mode = os.environ.get("SCGR_NVLS_P2P", "auto")
        if mode == "1":
            os.environ.setdefault("SCGR_NVLS_P2P_SHOTS", "3")      # forced: both shots, whatever the world size (read once by libscgr)
        return world in (2, 4, 8) and (mode == "1" or (mode == "auto" and world <= self.P2P_MAX_WORLD))

    def timed_out(self) -> bool:
        """True if a barrier inside the fused collective ever gave up waiting (the sums are then wrong)."""
        return self._sync is not None and bool(int(self._sync[2].item()) != 0)

    def _sparse(self) -> bool:
        return self.row_floats > 0 and os.environ.get("SCGR_ALLREDUCE_SPARSE", "1") != "0"

    def all_reduce(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """THE collective of the path: one SUM all-reduce over the flat buffer."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return None
        if self._symm is not None and group is None and not async_op:
            from . import _lib
            hdl = self._symm
            stream = C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
            lib = _lib.load()
            rank, world = int(hdl.rank), int(hdl.world_size)
            # multicast address of flat[0]: same offset from the multicast base as from this rank's own mapping
            mc = int(hdl.multicast_ptr) + (self.flat.data_ptr() - int(hdl.buffer_ptrs[rank]))
            if self._fused(world):
                # ONE launch: barrier -> dense shot -> barrier -> row-sparse shot -> barrier, the barriers inside the kernel
                sparse = self._sparse()
                f = _lib.ScgrNvlsFused(mc, self.dense_floats if sparse else self.flat.numel(),
                                       (mc + 4 * self.rows_offset) if sparse else None,
                                       self.views["live"].data_ptr() if sparse else None, self.P if sparse else 0,
                                       self.row_floats, rank, world)
                off = self.flat.data_ptr() - int(hdl.buffer_ptrs[rank]) + 4 * self._flag_offset
                for q in range(world):
                    f.flags[q] = int(hdl.buffer_ptrs[q]) + off
                f.sync_local = self._sync.data_ptr()
                f.epoch = self._epoch
                if self._p2p(world):
                    base_off = self.flat.data_ptr() - int(hdl.buffer_ptrs[rank])
                    for q in range(world):
                        f.peer_ptrs[q] = int(hdl.buffer_ptrs[q]) + base_off
                self._epoch = (self._epoch + 3) & 0xffffffff      # (compared modulo 2^32 in the kernel)
                _lib.check(lib.scgr_nvls_allreduce_fused(C.byref(f), stream))
                return None
            hdl.barrier(channel=0)           # every replica has been written by its rank's backward
            if not self._sparse():
                _lib.check(lib.scgr_nvls_allreduce(C.c_void_p(mc), self.flat.numel(), rank, world, stream))
            else:
                # shot A (dense): the small blocks, the statistics and the live counts
                _lib.check(lib.scgr_nvls_allreduce(C.c_void_p(mc), self.dense_floats, rank, world, stream))
                hdl.barrier(channel=1)       # the summed live counts are in place on every rank
                # shot B (row-sparse): dL/dSH rows of the Gaussians that are live on some rank
                _lib.check(lib.scgr_nvls_allreduce_rows(C.c_void_p(mc + 4 * self.rows_offset), self.views["live"].data_ptr(),
                                                        self.P, self.row_floats, rank, world, stream))
            hdl.barrier(channel=2)           # every shard has been broadcast
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    @property
    def payload(self) -> torch.Tensor:
        """The flat buffer without its tail padding (up to 4 * world - 1 floats that no field owns: the dense paths sum
        them, the row-sparse shot has no reason to)."""
        return self.flat[:self.payload_floats]

    def nbytes(self) -> int:
        return self.flat.numel() * 4


def shard_views(n_views: int, rank: int, world_size: int):
    """Views of a batch handled by `rank`: round-robin, one view per GPU when n_views == world_size."""
    return list(range(rank, n_views, world_size))
