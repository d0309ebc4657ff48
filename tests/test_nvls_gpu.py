"""The path's single collective on real GPUs (needs >= 2 of them: `gpurun --gpus 2|4|8`; skipped on a 1-GPU box).

libscgr's own two-shot NVSwitch all-reduce (scgaussian_b200/csrc/collective.cu: multimem.ld_reduce + multimem.st over a
symmetric-memory multicast mapping) must give, on every rank, the sum of the per-rank buffers: it is compared BIT FOR
BIT with ncclAllReduce on a private copy of the same data (dyadic values: every partial sum is exact in fp32, so any
summation order gives the same bits) -- for a sparse set of live Gaussians (the row-sparse dL/dSH shot skips the
rest), for none and for all of them; then once more on random data against NCCL within fp32 reassociation.  The buffer is the FlatGradBuffer the backward kernels write into, at BASELINE config 3's size."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, P, out_dir, fused):
    import torch.distributed as dist
    # fused: "1" one launch (shots chosen automatically), "0" two launches between host-issued barriers,
    #        "mc" one launch with the multicast shots, "p2p" one launch with the peer-to-peer shots
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SCGR_ALLREDUCE_FUSED="0" if fused == "0" else "1",
                      SCGR_NVLS_P2P={"mc": "0", "p2p": "1"}.get(fused, "auto"))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {"rank": rank}
    try:
        from scgaussian_b200.parallel import FlatGradBuffer
        try:
            buf = FlatGradBuffer(P, sh_coeffs=16, device=dev, symmetric=True)      # force the NVLS path at any world size
        except Exception as e:
            res["skip"] = f"symmetric memory / multicast unavailable: {e!r}"[:300]
            torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
            return
        res["collective"] = buf.collective
        n = buf.flat.numel()
        base = ((torch.arange(n, device=dev, dtype=torch.int64) % 1021) - 510).to(torch.float32) / 64.0

        def fill(values, density, seed):
            """A buffer as the backward leaves it: arbitrary small blocks, `live` = 1 on a random subset of the Gaussians,
            dL/dSH rows non-zero only there."""
            buf.flat.copy_(values)
            g = torch.Generator(device=dev).manual_seed(seed)
            live = (torch.rand(P, device=dev, generator=g) < density).to(torch.float32)
            buf.views["live"].copy_(live)
            buf.views["shs"].mul_(live.view(P, 1, 1))
            return buf.flat.clone()

        for rep, density in enumerate((0.15, 0.0, 1.0)):            # sparse, nobody live, everybody live
            mine = fill(base * (rank + 1 + rep), density, 1000 * rep + rank)
            buf.all_reduce()
            dist.all_reduce(mine)                                   # NCCL, dense, on a private copy of the same data
            # dyadic values: any summation order, same bits.  (Compared over the fields: with 12 P % world != 0 -- P = 1001 at
            # 8 ranks -- the buffer ends in padding that NCCL sums and the row-sparse shot rightly never touches.)
            npay = buf.payload_floats
            res[f"equals_nccl_{rep}"] = bool(torch.equal(buf.payload, mine[:npay]))
            res[f"live_max_{rep}"] = float(buf.views["live"].max())
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        mine = fill(torch.randn(n, device=dev, generator=g), 0.2, 77 + rank)
        buf.all_reduce()
        dist.all_reduce(mine)
        res["random_max_abs_diff_vs_nccl"] = float((buf.payload - mine[:buf.payload_floats]).abs().max())
        res["random_scale"] = float(mine.abs().max())
        # replicas must be identical on every rank (one reduction per element, then a broadcast)
        digest = buf.payload.double().sum().reshape(1)
        lo, hi = digest.clone(), digest.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["replicas_identical"] = bool(lo.item() == hi.item())
        torch.cuda.synchronize()
        res["timed_out"] = buf.timed_out()
    finally:
        torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", ["1", "0", "mc", "p2p"])      # see _worker
@pytest.mark.parametrize("P", [1_000_000, 1001])
def test_nvls_allreduce_equals_sum_of_rank_buffers(tmp_path, P, fused):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on one box (gpurun --gpus N)")
    import torch.multiprocessing as mp
    world = min(8, torch.cuda.device_count())
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, P, str(tmp_path), fused), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    if any("skip" in r for r in res):
        pytest.skip(next(r["skip"] for r in res if "skip" in r))
    for r in res:
        assert r["collective"].startswith("nvls"), r
        for rep in range(3):
            assert r[f"equals_nccl_{rep}"], r
        assert r["live_max_0"] >= 1.0 and r["live_max_1"] == 0.0 and r["live_max_2"] == float(world), r
        assert r["random_max_abs_diff_vs_nccl"] <= 1e-5 * r["random_scale"], r
        assert r["replicas_identical"], r
        assert not r["timed_out"], r
        assert ("one launch" in r["collective"]) == (fused != "0"), r
        if fused in ("mc", "p2p"):
            assert ("peer-to-peer" in r["collective"]) == (fused == "p2p"), r
    try:
        import json
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": "nvls_allreduce", "world": world, "P": P, "fused": fused, "ranks": res}) + "\n")
    except Exception:
        pass
