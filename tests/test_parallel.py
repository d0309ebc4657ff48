"""CPU (gloo, world_size 2) test of the view-sharded data-parallel host logic
(scgaussian_b200/parallel.py): the single flat all-reduce must equal the sum of the per-view
gradients, including the densification statistics.  The per-view compute here is the oracle --
it stands in for the CUDA rasterizer, which needs a GPU; the collective plumbing is what is tested."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scgaussian_b200.parallel import FlatGradBuffer, shard_views
from tests import util
from oracle import torch_oracle as O

P, W, H = 120, 48, 32


def _view_grads(view_id):
    case = util.make_case(P, W, H, sh_degree=1, max_sh_degree=1, scale_median=0.08,
                          w2c=O.yaw_w2c((view_id - 0.5) * 4.0))
    gC, gD, gA = O.synth_upstream_grads(W, H, seed=10 + view_id)
    co, (c, radii, d, a), g = util.run_c_oracle(case, "f32", grads=(gC, gD, gA))
    return radii, g


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        buf = FlatGradBuffer(P, sh_coeffs=4, device="cpu")
        assert shard_views(world, rank, world) == [rank]
        radii, g = _view_grads(rank)
        out = buf.out_dict()
        for k in out:                      # what scgr_backward does on the GPU: fill the views in place
            if k == "live":                # ScgrGrads.live_count: 1 where the view gave the Gaussian any gradient
                out[k].copy_(torch.from_numpy((g["opacities"][:, 0] != 0).astype("float32")))
            elif k != "stats":
                out[k].copy_(torch.from_numpy(g[k]).reshape(out[k].shape))
        buf.fill_stats(torch.from_numpy(radii))
        buf.all_reduce()
        np.save(os.path.join(out_dir, f"flat_{rank}.npy"), buf.flat.numpy())
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_equals_sum_of_view_gradients(tmp_path):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    flats = [np.load(tmp_path / f"flat_{r}.npy") for r in range(world)]
    assert np.array_equal(flats[0], flats[1])            # replicas stay bit-identical
    ref = FlatGradBuffer(P, sh_coeffs=4, device="cpu")
    ref.flat.zero_()
    for v in range(world):
        radii, g = _view_grads(v)
        for k, t in ref.views.items():
            if k == "stats":
                vis = torch.from_numpy(radii > 0).float()
                t[:, 0] += torch.linalg.vector_norm(torch.from_numpy(g["means2D"])[:, :2], dim=-1) * vis
                t[:, 1] += vis
            elif k == "live":
                t += torch.from_numpy((g["opacities"][:, 0] != 0).astype("float32"))
            else:
                t += torch.from_numpy(g[k]).reshape(t.shape)
    np.testing.assert_allclose(flats[0], ref.flat.numpy(), rtol=1e-6, atol=1e-9)
    assert ref.views["stats"][:, 1].max() <= world


def test_flat_buffer_layout():
    b = FlatGradBuffer(10, sh_coeffs=16, device="cpu")
    assert [n for n, _ in b.fields] == ["means3D", "opacities", "scales", "rotations", "stats", "live", "shs"]
    assert b.flat.numel() >= 10 * (3 + 48 + 1 + 3 + 4 + 2 + 1)
    assert b.dense_floats % 4 == 0 and b.rows_offset == b.dense_floats and b.row_floats == 48     # the row-sparse shot's block
    for v in b.views.values():
        assert v.is_contiguous() and v.data_ptr() % 16 == 0
    b2 = FlatGradBuffer(10, use_sh=False, use_cov=True, device="cpu", with_stats=False)
    assert [n for n, _ in b2.fields] == ["means3D", "colors_precomp", "opacities", "cov3D_precomp", "live"]
    assert b2.row_floats == 0                     # no SH block: the whole buffer goes through the dense shot
