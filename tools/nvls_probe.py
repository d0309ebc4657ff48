"""Multi-GPU probe (torchrun, GPU box): the NVLS two-shot all-reduce of libscgr against ncclAllReduce on the
flat gradient buffer of config 3 (244 MB): bit-level agreement of the replicas, agreement with NCCL, time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from scgaussian_b200.parallel import FlatGradBuffer

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
buf = FlatGradBuffer(P, sh_coeffs=16, device=dev, symmetric=True)
if rank == 0:
    print("collective:", buf.collective, getattr(buf, "_symm_error", ""), "numel", buf.flat.numel(), flush=True)
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = torch.randn(buf.flat.numel(), device=dev, generator=g)
ref = src.clone()
dist.all_reduce(ref)
buf.flat.copy_(src)
buf.all_reduce()
torch.cuda.synchronize()
err = float((buf.flat - ref).abs().max()) / float(ref.abs().max())
chk = buf.flat.double().sum().reshape(1)
all_chk = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(all_chk, chk)
same = all(float(c) == float(all_chk[0]) for c in all_chk)
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_ours = timeit(lambda: buf.all_reduce())
t_nccl = timeit(lambda: dist.all_reduce(ref))
if rank == 0:
    print(f"world {world}: max rel err vs nccl {err:.2e}, replicas identical {same}, "
          f"ours {t_ours * 1000:.0f} us, nccl {t_nccl * 1000:.0f} us for {buf.nbytes() / 1e6:.0f} MB", flush=True)
dist.destroy_process_group()
