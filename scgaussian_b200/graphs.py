"""CUDA-graph replay of a whole small-scene training step (SURVEY.md section 7 hard part 5).

At the reference's own training resolution (504x378, reference README.md:65 `-r 8`) one forward + loss + backward is
~260 us of GPU work against ~400 us of Python / autograd / launch overhead on the host: the step is host-bound however
fast the kernels are.  Everything the operator enqueues is capturable once the forward does not wait for the instance
count on the host -- the `async` binning protocol of scgaussian_b200/rasterizer.py (both stages enqueued blind with a
pre-sized buffer; every kernel reads R on the device and refuses to run past the buffer) -- so the whole step can be
recorded ONCE with torch.cuda.CUDAGraph and replayed with a single launch:

    step = GraphedStep(lambda: loss_fn(render(cam_static, gaussians, pipe, bg)).backward())
    for it in range(n):
        cam_static.world_view_transform.copy_(...)   # per-step inputs are written INTO the captured tensors
        step.replay()                                # forward + loss + backward: one cudaGraphLaunch
        optimizer.step()                             # .grad tensors are the captured ones, filled by the replay

Rules (the usual ones of whole-network capture): every tensor the closure reads must keep its address (update with
copy_), shapes are fixed -- P, W, H, SH degree, and the binning capacity chosen at capture time (ASYNC_HEADROOM x the
instance count of the warm-up views; a replayed view that outgrows it is dropped: NaN images, zero gradients, and
`overflowed()` says so) --, and gradients accumulate into the tensors allocated during capture (`zero_grad(set_to_none
=False)` or overwrite semantics: the closure should not depend on .grad being None).
"""
from __future__ import annotations

from typing import Callable

import torch

from . import rasterizer as R


class GraphedStep:
    def __init__(self, fn: Callable[[], None], warmup: int = 3, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._dev = dev
        self._mode = R._BINNING_MODE
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        R._BINNING_MODE = "async"
        try:
            with torch.cuda.stream(side):
                for _ in range(max(2, warmup)):        # first call learns R (blocking protocol), the rest run async
                    fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)                # the pinned status words now hold the warm-up views' counts
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                fn()
        finally:
            R._BINNING_MODE = self._mode
        self._status = R._status_buffer(dev)

    def replay(self) -> None:
        self.graph.replay()

    def overflowed(self) -> bool:
        """True when a replayed view did not fit the captured binning buffer (its images are NaN, its gradients zero).
        Reads the pinned status words the replay's own copy nodes fill: call after a synchronisation point."""
        return bool(int(self._status[3]) != 0)

    def num_rendered(self) -> int:
        """Instance count of the last completed replay (same caveat as overflowed())."""
        return int(self._status[2])
