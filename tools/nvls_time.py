"""Times the path's collective alone (torchrun, N GPUs of one box): FlatGradBuffer.all_reduce() at BASELINE config 3's
size for several live densities, fused (one launch, in-kernel barriers) and unfused, against dist.all_reduce (NCCL) on a
plain copy.  CUDA events around 20 back-to-back calls after warm-up, max over ranks.  Development tool.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/nvls_time.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from scgaussian_b200.parallel import FlatGradBuffer
    P = int(os.environ.get("P", 1_000_000))
    out = {"world": world, "P": P}

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for fused in ("1", "p2p", "0"):      # one launch (multicast shots), one launch (peer-to-peer shots), two launches
        os.environ["SCGR_ALLREDUCE_FUSED"] = "0" if fused == "0" else "1"
        os.environ["SCGR_NVLS_P2P"] = "1" if fused == "p2p" else "0"
        for sparse in ("1", "0"):
            os.environ["SCGR_ALLREDUCE_SPARSE"] = sparse
            buf = FlatGradBuffer(P, sh_coeffs=16, device=dev, symmetric=True)
            for density in ((0.0, 0.45, 1.0) if sparse == "1" else (1.0,)):
                g = torch.Generator(device=dev).manual_seed(7 + rank)
                buf.flat.zero_()
                live = (torch.rand(P, device=dev, generator=g) < density).float()
                buf.views["live"].copy_(live)
                out[f"fused{fused}_sparse{sparse}_live{density}"] = timed(buf.all_reduce)
            assert not buf.timed_out()
            del buf
    plain = torch.zeros(62 * P, device=dev)
    out["nccl_dense"] = timed(lambda: dist.all_reduce(plain))
    small = torch.zeros(14 * P, device=dev)
    out["nccl_small_blocks_only"] = timed(lambda: dist.all_reduce(small))
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"nvls_time_{world}.json"), "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
