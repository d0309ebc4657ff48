#!/usr/bin/env python
"""bench.py -- views/sec forward+backward of the rasterizer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 3 -- synthetic 1M Gaussians, 1920x1080, SH degree 3, seed-0 generator of
SURVEY.md section 8d.  A *step* is one forward + backward of one view per GPU (weak scaling) followed, when N > 1, by
the path's single collective: one all-reduce over the flat gradient buffer.  The camera changes EVERY step -- the 8
yawed cameras of SURVEY 8d, (k - 3.5) * 2 degrees; rank r renders camera (r + step) % 8 -- so the instance count R varies
from step to step the way it does in training (reference train.py:135 picks a random camera per iteration) and the
pre-sized binning buffer can miss (`need_capacity_hits`).

Printed JSON line (rank 0):
  value          views/s over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            the same metric through the public operator (GaussianRasterizer + autograd), with the per-step host
                 inputs of a training step (camera + ground-truth image, pinned host memory -> device) and the loss
                 read back (device -> host) inside the timed region
  roofline       dominant kernel: algorithmic bytes / its CUDA-event duration
  roofline_issue the render kernels against the SM issue rate (148 SMs x 4 schedulers x f_clk) and per pair evaluation
  kernels        per-kernel average milliseconds of one step (profiled in a separate short loop)
  batch8         the north_star's fixed 8-view batch: 8 / N views per rank, gradients accumulated on the device,
                 ONE all-reduce per batch
  config4        (N = 8, or --config4) BASELINE config 4: 5M Gaussians, 3840x2160, one view per GPU, 1.22 GB all-reduce
  gpu_standin_baseline  a deliberately naive CUDA restatement of the published algorithm (baseline/standin: one
                 thread per Gaussian, cub 64-bit radix sort, 256-thread tiles, 10 atomics per pair) on the same
                 inputs on the same GPU -- NOT the reference, which is not obtainable offline
  cpu_baseline   the CPU oracle (scalar C port, OpenMP) timed on this host on full views
  --impl reference: the reference has no CPU path and its CUDA rasterizer is not obtainable offline (SURVEY.md
  section 0); this arm times the oracle port of its algorithm on all host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    3: dict(P=1_000_000, W=1920, H=1080, sh=3, scale=0.01,
            name="config3: synthetic 1M Gaussians, 1920x1080, SH deg 3, seed 0 (SURVEY 8d)"),
    4: dict(P=5_000_000, W=3840, H=2160, sh=3, scale=0.005,
            name="config4: synthetic 5M Gaussians, 3840x2160, SH deg 3, seed 0 (SURVEY 8d)"),
}
N_CAMERAS = 8
METRIC = "views/sec fwd+bwd @ 1M Gaussians, 1080p, SH3"
UNIT = "views/s"


def camera_yaw(k: int) -> float:
    return (k - 3.5) * 2.0          # the 8 cameras of a batch


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler runs from
    the start of the warm-up (the GPU is loaded without a gap from there to the end of the timed region); every row is
    stamped on arrival, and the rows that fall inside the timed region are the ones reported.  The timed region of the
    default run lasts a few tens of milliseconds -- shorter than nvidia-smi's sampling period -- so when no row landed in
    it the rows taken under the same uninterrupted load just before it are used, and `window` says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t_load = self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def under_load(self):          # the warm-up has started: rows from here on see a busy GPU
        self.t_load = time.perf_counter()

    def n_under_load(self):
        return sum(1 for t, _ in self.rows if self.t_load is not None and t >= self.t_load + 0.05)

    def timed(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = list(self.rows)
        inside = [r for t, r in rows if self.t0 is not None and self.t0 <= t <= self.t1]
        window = "timed region"
        if not inside:
            inside = [r for t, r in rows if self.t_load is not None and self.t_load + 0.05 <= t <= (self.t1 or t)]
            window = "warm-up + timed region (one uninterrupted load; the timed region is shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: the CPU oracle port on all host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_views_per_s(reps: int, cfg: dict):
    """Times the scalar C oracle (all host cores, OpenMP) on full views of the workload, cycling the cameras."""
    from oracle import torch_oracle as O
    from oracle.c_oracle import COracle
    sc = O.synth_scene(cfg["P"], cfg["W"], cfg["H"], sh_degree=cfg["sh"], scale_median=cfg["scale"], seed=0)
    gC, gD, gA = [g.numpy() for g in O.synth_upstream_grads(cfg["W"], cfg["H"], seed=1)]
    cores = os.cpu_count() or 1
    co = COracle("f32")
    co.set_threads(cores)          # torchrun exports OMP_NUM_THREADS=1: ask OpenMP for every core explicitly
    times = []
    for i in range(reps):
        cam = O.make_camera(cfg["W"], cfg["H"], w2c=O.yaw_w2c(camera_yaw(i % N_CAMERAS)))
        kw = dict(means3D=sc["means3D"].numpy(), opacities=sc["opacities"].numpy(), W=cfg["W"], H=cfg["H"],
                  tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=[0.0, 0.0, 0.0],
                  viewmatrix=cam["viewmatrix"].numpy(), projmatrix=cam["projmatrix"].numpy(),
                  campos=cam["campos"].numpy(), sh_degree=cfg["sh"], shs=sc["shs"].numpy(),
                  scales=sc["scales"].numpy(), rotations=sc["rotations"].numpy())
        t0 = time.perf_counter()
        co.forward(**kw)
        co.backward(gC, gD, gA)
        times.append(time.perf_counter() - t0)
    return times, cores


def run_reference_arm(args, rank, cfg):
    """--impl reference: CPU oracle port of the reference's algorithm (the reference itself has no CPU path and its
    CUDA package is not in /root/reference nor installable offline).  Rank 0 alone runs it."""
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    steps = max(1, args.steps)
    warm = max(0, min(args.warmup, 1))       # each step is ~seconds of CPU work: one warm-up is plenty
    times, cores = cpu_oracle_views_per_s(warm + steps, cfg)
    times = times[warm:]
    ms = 1000.0 * sum(times) / len(times)
    v = 1000.0 / ms
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "note": "reference CUDA rasterizer unavailable offline; this is the CPU oracle port of its algorithm"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} full view(s) fwd+bwd ({cfg['P']} Gaussians, {cfg['W']}x{cfg['H']}), OpenMP over {cores} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Workload:
    """Scene + the 8 cameras + upstream gradients + the flat gradient buffer of one config, resident on `dev`."""

    def __init__(self, cfg, dev, rank, world):
        from scgaussian_b200 import GaussianRasterizationSettings, synthetic as O
        from scgaussian_b200.parallel import FlatGradBuffer
        self.cfg, self.dev, self.rank, self.world = cfg, dev, rank, world
        P, W, H = cfg["P"], cfg["W"], cfg["H"]
        sc = O.synth_scene(P, W, H, sh_degree=cfg["sh"], scale_median=cfg["scale"], seed=0)
        self.scene_cpu = sc
        self.t = {k: v.to(dev).contiguous() for k, v in sc.items()}
        self.grads = [g.to(dev).contiguous() for g in O.synth_upstream_grads(W, H, seed=1)]
        self.cams_cpu = [O.make_camera(W, H, w2c=O.yaw_w2c(camera_yaw(k))) for k in range(N_CAMERAS)]
        bg = torch.zeros(3, device=dev)
        self.settings = [GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=c["tanfovx"], tanfovy=c["tanfovy"], bg=bg, scale_modifier=1.0,
            viewmatrix=c["viewmatrix"].to(dev), projmatrix=c["projmatrix"].to(dev), sh_degree=cfg["sh"],
            campos=c["campos"].to(dev), prefiltered=False, debug=False) for c in self.cams_cpu]
        self.args_in = (self.t["means3D"], self.t["opacities"], self.t["shs"], None, self.t["scales"], self.t["rotations"], None)
        self.flat = FlatGradBuffer(P, sh_coeffs=(cfg["sh"] + 1) ** 2, device=dev)
        self.out_views = self.flat.out_dict()
        self.R_seen = []

    def view(self, i):
        """Forward + backward of camera i (mod 8) into the flat buffer; returns the ForwardState."""
        from scgaussian_b200 import rasterizer as R
        s = self.settings[i % N_CAMERAS]
        color, radii, depth, alpha, state = R.rasterize_forward_raw(*self.args_in, s)
        R.rasterize_backward_raw(state, *self.args_in, s, *self.grads, out=self.out_views)
        return state

    def step(self, i, collective=True):
        state = self.view(self.rank + i)
        if self.world > 1 and collective:
            self.flat.all_reduce()               # THE collective of the path (statistics ride in the same buffer)
        self.R_seen.append(int(state.num_rendered))     # host int already read by the forward: free
        return state

    def batch(self, i, views_per_rank):
        """north_star's 8-view batch: this rank renders views_per_rank views, accumulating on the device; one all-reduce."""
        from scgaussian_b200 import rasterizer as R
        for v in range(views_per_rank):
            s = self.settings[(self.rank * views_per_rank + v + i) % N_CAMERAS]
            color, radii, depth, alpha, state = R.rasterize_forward_raw(*self.args_in, s)
            R.rasterize_backward_raw(state, *self.args_in, s, *self.grads, out=self.out_views, accumulate=v > 0)
        if self.world > 1:
            self.flat.all_reduce()


def timed(fn, n, barrier, dev, world):
    """n calls of fn(i) bracketed by barrier + synchronize on both sides; CUDA events; max over ranks.  ms per call."""
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4], help="workload of the headline numbers (default: BASELINE config 3)")
    ap.add_argument("--config4", action="store_true", help="add the config4 entry at any N (default: only at N = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--no-standin", action="store_true")
    ap.add_argument("--no-batch8", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[args.config]

    if args.impl == "reference":
        run_reference_arm(args, rank, cfg)
        return

    import torch.distributed as dist
    from scgaussian_b200 import GaussianRasterizer, _lib
    from scgaussian_b200 import rasterizer as R
    from scgaussian_b200.losses import photometric_loss

    lib = _lib.load()                      # raises if libscgr.so is missing: no fallback
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    W_ = max(3, args.warmup)
    K = max(1, args.steps)
    P_GAUSS, WIDTH, HEIGHT, SH_DEG = cfg["P"], cfg["W"], cfg["H"], cfg["sh"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = Workload(cfg, dev, rank, world)
    flat = wl.flat

    # ---------------- the collective's result, checked once outside the timed region ----------------
    allreduce_check = None
    if world > 1:
        n = flat.flat.numel()
        # small dyadic values: every partial sum is exact in fp32, so the result must equal the closed form BIT FOR BIT
        # whatever order the switch / the ring adds in
        base = ((torch.arange(n, device=dev, dtype=torch.int64) % 1021) - 510).to(torch.float32) / 64.0
        flat.flat.copy_(base * (rank + 1))
        if "live" in flat.views and "shs" in flat.views:
            # the buffer as the backward leaves it: `live` = 1 on a (rank-dependent) subset of the Gaussians, dL/dSH rows
            # non-zero only there -- the contract the row-sparse shot relies on
            gi = torch.arange(P_GAUSS, device=dev)
            live = ((gi + 3 * rank) % 7 < 2).to(torch.float32)
            flat.views["live"].copy_(live)
            flat.views["shs"].mul_(live.view(-1, 1, 1))
        mine = flat.flat.clone()
        dist.barrier()                       # (the in-kernel barriers of the fused collective wait ~10 s at most)
        flat.all_reduce()
        got = flat.flat.clone()
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)                       # NCCL on a private copy of the same data
        # dyadic values: any summation order gives the same bits (over the fields: the tail padding belongs to no field)
        err_nccl = float((got - mine)[:flat.payload_floats].abs().max())
        sm = flat.views["means3D"]                         # a dense block against its closed form
        off = (sm.data_ptr() - flat.flat.data_ptr()) // 4
        want = base[off:off + sm.numel()] * (world * (world + 1) // 2)
        err_exact = float((got[off:off + sm.numel()] - want).abs().max())
        ok = err_exact == 0.0 and err_nccl == 0.0
        okt = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        allreduce_check = {"ok": bool(okt.item() == 1.0), "collective": flat.collective, "n_floats": n,
                           "max_abs_diff_vs_nccl": err_nccl, "max_abs_diff_vs_closed_form": err_exact}
        del base, mine, got, want
        flat.flat.zero_()
        assert allreduce_check["ok"], allreduce_check

    # ---------------- device-resident throughput (value) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)               # nvidia-smi is up and printing before the load starts
        sampler.under_load()
    for i in range(W_):
        wl.step(i)
    # extra untimed steps (same on every rank) until the clock sampler has seen the loaded GPU a few times: the timed
    # region follows without a gap, so these rows describe its clocks even if none lands inside it
    settle = 0
    while True:
        go = torch.tensor([1 if (rank == 0 and sampler.proc is not None and sampler.n_under_load() < 4 and settle < 2000) else 0], device=dev)
        if world > 1:
            dist.broadcast(go, 0)
        if int(go.item()) == 0:
            break
        for i in range(16):
            wl.step(W_ + settle + i)
        settle += 16
    barrier()
    wl.R_seen.clear()
    l0 = lib.scgr_kernel_launch_count()
    nc0 = R.need_capacity_count
    t_0 = time.perf_counter()
    ms_step = timed(lambda i: wl.step(W_ + settle + i), K, barrier, dev, world)
    t_1 = time.perf_counter()
    launches = lib.scgr_kernel_launch_count() - l0
    need_capacity_hits = R.need_capacity_count - nc0
    if allreduce_check is not None:
        to = torch.tensor([1.0 if flat.timed_out() else 0.0], device=dev)
        dist.all_reduce(to, op=dist.ReduceOp.MAX)
        allreduce_check["a_barrier_timed_out_during_the_run"] = bool(to.item() != 0)
        assert not allreduce_check["a_barrier_timed_out_during_the_run"], allreduce_check
    clocks = None
    if rank == 0:
        sampler.timed(t_0, t_1)
        clocks = sampler.stop()
        clocks["extra_untimed_steps_before_timing"] = settle
    value = world * 1000.0 / ms_step
    R_list = list(wl.R_seen)
    R_mean = sum(R_list) / len(R_list)

    # ---------------- per-kernel profile (separate short loop; not the reported value) ----------------
    kernels = {}
    if rank == 0:
        lib.scgr_profile_enable(1)
        nprof = N_CAMERAS
        wl.R_seen.clear()
        for i in range(nprof):
            wl.step(i, collective=False)          # rank 0 only: must not enter a collective
        torch.cuda.synchronize()
        R_prof = sum(wl.R_seen) / len(wl.R_seen)
        import ctypes as C
        names = (C.c_char_p * 8192)()
        msarr = (C.c_float * 8192)()
        n = lib.scgr_profile_fetch(names, msarr, 8192)
        lib.scgr_profile_enable(0)
        for i in range(max(n, 0)):
            kernels.setdefault(names[i].decode(), []).append(float(msarr[i]))
        kernels = {k: {"ms_per_step": sum(v) / nprof, "launches_per_step": len(v) / nprof,
                       "ms_per_launch": sum(v) / len(v)} for k, v in kernels.items()}
        # Gaussians that receive any gradient in one view: what preprocess_backward actually has to read / compute
        live_frac = float((flat.views["opacities"] != 0).float().mean())
    if world > 1:
        barrier()

    # ---------------- the fixed 8-view batch (north_star): 8 / N views per rank, one all-reduce per batch ----------------
    batch8 = None
    if not args.no_batch8 and N_CAMERAS % world == 0:
        vpr = N_CAMERAS // world
        for i in range(2):
            wl.batch(i, vpr)
        nb = max(2, K // 4)
        ms_b = timed(lambda i: wl.batch(i, vpr), nb, barrier, dev, world)
        batch8 = {"views_per_batch": N_CAMERAS, "views_per_rank": vpr, "ms_per_batch": ms_b,
                  "views_s": N_CAMERAS * 1000.0 / ms_b, "batches": nb, "scaling": "strong (fixed 8-view batch)",
                  "what": "each rank renders 8/N views, scgr_backward accumulates on the device (ScgrGrads.accumulate), "
                          "one all-reduce of the flat buffer per batch"}

    # ---------------- end-to-end through the public operator (e2e) ----------------
    e2e = None
    if not args.no_e2e:
        # per-step host inputs of a training step (reference train.py:135-147): the camera (view / projection
        # matrices, camera centre, background: 38 floats, ONE packed pinned buffer -> one H2D copy) and the
        # ground-truth image (pinned -> device on a copy stream).  Two device buffer pairs: the uploads of step k+1
        # are issued while step k is still running (what a prefetching data loader does), and the loss of step k
        # is read back (D2H into pinned memory + event) after step k+1 has been enqueued, so the GPU never waits
        # for Python.  Every step's uploads and every step's loss read-back happen inside the timed region;
        # nothing is cached across steps; the camera changes every step.
        gt_host = [torch.rand(3, HEIGHT, WIDTH).pin_memory() for _ in range(2)]
        cam_host = [torch.cat([c["viewmatrix"].reshape(-1), c["projmatrix"].reshape(-1), c["campos"].reshape(-1),
                               torch.zeros(3)]).contiguous().pin_memory() for c in wl.cams_cpu]
        leaves = {k: wl.t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        h2d = gt_host[0].numel() * 4 + cam_host[0].numel() * 4
        loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        upload_done = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        gt_dev = [torch.empty(3, HEIGHT, WIDTH, device=dev) for _ in range(2)]
        cam_dev = [torch.empty(38, device=dev) for _ in range(2)]
        m2d = torch.zeros(P_GAUSS, 3, device=dev, requires_grad=True)   # never read by the op; only its .grad matters
        losses = []
        s0 = wl.settings[0]

        def upload(k):
            b = k & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # the step that last used this buffer pair is done with it
                cam_dev[b].copy_(cam_host[(rank + k) % N_CAMERAS], non_blocking=True)
                gt_dev[b].copy_(gt_host[b], non_blocking=True)
                upload_done[b].record(copy_stream)

        def launch(k):
            b = k & 1
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(upload_done[b])
            c = cam_dev[b]
            s2 = s0._replace(viewmatrix=c[0:16].view(4, 4), projmatrix=c[16:32].view(4, 4), campos=c[32:35], bg=c[35:38])
            color, radii, depth, alpha = GaussianRasterizer(s2)(
                means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                scales=leaves["scales"], rotations=leaves["rotations"])
            # the reference's training loss (train.py:160-161: L1 + 0.2 D-SSIM, fused: scgaussian_b200/losses.py)
            # plus terms that send gradient into the depth and alpha outputs (train.py:164-168 use both)
            loss = photometric_loss(color, gt_dev[b], 0.2) + (depth.sum() + alpha.sum()) * (0.01 / (HEIGHT * WIDTH))
            if world > 1:
                # one all-reduce for the public-API path too: the operator's autograd node writes gradients, statistics and
                # live counts straight into the flat buffer (FlatGradBuffer.capture), no packing copies
                with flat.capture():
                    loss.backward()
                flat.all_reduce()
            else:
                loss.backward()
            m2d.grad = None
            consumed[b].record(cur)
            loss_host[b].copy_(loss.detach().reshape(1), non_blocking=True)   # D2H read of the step's result
            loss_ready[b].record(cur)
            for v in leaves.values():
                v.grad = None

        def read_loss(k):
            loss_ready[k & 1].synchronize()
            losses.append(float(loss_host[k & 1][0]))

        def e2e_run(n):
            for b in range(2):
                consumed[b].record(torch.cuda.current_stream(dev))
            upload(0)
            for k in range(n):
                if k + 1 < n:
                    upload(k + 1)
                launch(k)
                if k > 0:
                    read_loss(k - 1)
            read_loss(n - 1)

        e2e_run(W_)
        barrier()
        ke = max(N_CAMERAS, K // 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        e2e_run(ke)
        e1.record()
        barrier()
        assert len(losses) == W_ + ke and all(math.isfinite(x) for x in losses)
        ms_e = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * 1000.0 * ke / float(ms_e.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "steps": ke,
               "what": "GaussianRasterizer + the reference's L1 + 0.2 D-SSIM loss (fused) + depth/alpha means + autograd backward through the public operators; every step uploads "
                       "its camera (38 floats, one packed copy; a different camera every step) and ground-truth image from pinned host memory and reads its loss "
                       "back; uploads of step k+1 are prefetched on a copy stream while step k runs and the loss of step k is "
                       "read (pinned + event) after step k+1 is enqueued -- all inside the timed region; Gaussian parameters "
                       "stay resident (they are model state, reference train.py)"}
        # ---- the same end-to-end step replayed as ONE CUDA graph (scgaussian_b200/graphs.py; N = 1): extra entry ----
        # forward + loss + backward through the public operator captured once under the never-blocking binning protocol;
        # every step still uploads its camera and ground-truth image from pinned memory (copy stream, one step ahead;
        # the replay reads them through a 25 MB device-to-device copy into the captured tensors) and reads its loss back.
        if world == 1:
            try:
                from scgaussian_b200.graphs import GraphedStep
                cam_static = torch.empty(38, device=dev)
                gt_static = torch.empty(3, HEIGHT, WIDTH, device=dev)
                loss_static = torch.zeros(1, device=dev)
                cam_static.copy_(cam_host[0])
                gt_static.copy_(gt_host[0])
                sg = s0._replace(viewmatrix=cam_static[0:16].view(4, 4), projmatrix=cam_static[16:32].view(4, 4),
                                 campos=cam_static[32:35], bg=cam_static[35:38])

                def graph_body():
                    for v in leaves.values():
                        v.grad = None                  # fresh gradient tensors from the graph's pool: no accumulate pass
                    m2d.grad = None
                    color, radii, depth, alpha = GaussianRasterizer(sg)(
                        means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                        scales=leaves["scales"], rotations=leaves["rotations"])
                    loss = photometric_loss(color, gt_static, 0.2) + (depth.sum() + alpha.sum()) * (0.01 / (HEIGHT * WIDTH))
                    loss.backward()
                    loss_static.copy_(loss.detach().reshape(1))

                gstep = GraphedStep(graph_body, warmup=3, device=dev)
                glosses = []

                def g_launch(k):
                    b = k & 1
                    cur = torch.cuda.current_stream(dev)
                    cur.wait_event(upload_done[b])
                    cam_static.copy_(cam_dev[b], non_blocking=True)
                    gt_static.copy_(gt_dev[b], non_blocking=True)
                    consumed[b].record(cur)
                    gstep.replay()
                    loss_host[b].copy_(loss_static, non_blocking=True)
                    loss_ready[b].record(cur)

                def g_run(n):
                    for b in range(2):
                        consumed[b].record(torch.cuda.current_stream(dev))
                    upload(0)
                    for k in range(n):
                        if k + 1 < n:
                            upload(k + 1)
                        g_launch(k)
                        if k > 0:
                            loss_ready[(k - 1) & 1].synchronize()
                            glosses.append(float(loss_host[(k - 1) & 1][0]))
                    loss_ready[(n - 1) & 1].synchronize()
                    glosses.append(float(loss_host[(n - 1) & 1][0]))

                g_run(W_)
                torch.cuda.synchronize()
                kg = max(2 * N_CAMERAS, K)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                g_run(kg)
                e1.record()
                torch.cuda.synchronize()
                ms_g = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)
                # same losses as the eager loop for the same (camera, image) pairs?
                ref_l = {}
                for k, x in enumerate(losses[W_:]):
                    ref_l.setdefault((k % N_CAMERAS, k & 1), x)
                diffs = [abs(x - ref_l[(k % N_CAMERAS, k & 1)]) for k, x in enumerate(glosses[W_:]) if (k % N_CAMERAS, k & 1) in ref_l]
                e2e["graph_replay"] = {"value": 1000.0 * kg / ms_g, "unit": UNIT, "steps": kg, "overflowed": gstep.overflowed(),
                                       "all_losses_finite": all(math.isfinite(x) for x in glosses),
                                       "max_abs_loss_diff_vs_eager": max(diffs) if diffs else None,
                                       "what": "the same step (public operator + loss + autograd, per-step uploads and loss read-back) "
                                               "captured once and replayed with one cudaGraphLaunch per step (GraphedStep); extra entry, "
                                               "the headline e2e value above is the eager loop"}
                del gstep
            except Exception as e:         # pragma: no cover  (an extra entry must never take the metric down)
                e2e["graph_replay"] = {"error": repr(e)[:300]}
        del leaves, gt_dev, m2d
        torch.cuda.empty_cache()

    # ---------------- BASELINE config 4: 5M Gaussians, 3840x2160, one view per GPU, 1.22 GB all-reduce ----------------
    config4 = None
    if (world == N_CAMERAS or args.config4) and args.config != 4:
        try:
            c4 = CONFIGS[4]
            del wl.t, wl.args_in, wl.out_views
            w3_flat = wl.flat
            wl.flat = None
            del flat, w3_flat
            torch.cuda.empty_cache()
            w4 = Workload(c4, dev, rank, world)
            for i in range(3):
                w4.step(i)
            w4.R_seen.clear()
            k4 = max(4, K // 3)
            ms4 = timed(lambda i: w4.step(3 + i), k4, barrier, dev, world)
            ar_ms = None
            if world > 1:
                ar_ms = timed(lambda i: w4.flat.all_reduce(), 5, barrier, dev, world)
            config4 = {"workload": c4["name"], "views_per_step": world, "ms_per_step": ms4, "views_s": world * 1000.0 / ms4,
                       "steps": k4, "num_rendered_R_this_rank": w4.R_seen, "allreduce_ms": ar_ms,
                       "grad_allreduce_bytes": w4.flat.nbytes() if world > 1 else 0, "collective": w4.flat.collective if world > 1 else "none (1 GPU)",
                       "cameras": "rank r renders camera (r + step) % 8, yaw (k - 3.5) * 2 deg"}
            del w4
            torch.cuda.empty_cache()
        except Exception as e:             # pragma: no cover  (an extra entry must never take the metric down)
            config4 = {"error": repr(e)[:300]}
        if world > 1:
            barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- the whole training step (SURVEY 8f rows f1-f3 around the operator; N = 1 only) ----------------
    # Not the metric: an extra entry, measured after the timed regions above.  reference train.py:135-208 at the same
    # size -- raw parameters -> activations + assembly -> operator -> L1 + D-SSIM -> backward -> statistics -> Adam --
    # with the fused passes of this repo and with the reference's chain of torch ops, same rasterizer and loss kernels.
    train_step = None
    if world == 1 and not args.no_train_step and args.config == 3:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("scgr_train_step_time", os.path.join(ROOT, "tools", "train_step_time.py"))
            tst = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(tst)
            torch.cuda.empty_cache()
            train_step = tst.measure(wl.scene_cpu, wl.cams_cpu[3], dev, WIDTH, HEIGHT, steps=10, warm=3)
            pk, _ = peaks()
            ku = train_step.get("fused", {}).get("kernel_us_per_step", {})
            n_ray, n_bg = train_step["n_ray"], train_step["n_bg"]
            # csrc/model.cu: algorithmic bytes per launch.  Default (split SH layout): the assembly kernels move the small
            # arrays only (ray set 60 B raw in, free set 44 B; 44 B out; backward 44 grads + raw in, 36 / 44 out)
            alg = {"assemble_forward": n_ray * 104 + n_bg * 88, "assemble_backward": n_ray * (44 + 60 + 36) + n_bg * (44 + 44 + 44),
                   "adam": (n_ray * 57 + n_bg * 59) * 28}
            train_step["hbm_frac"] = {k: alg[k] / (ku[k] * 1e-6) / 1e9 / pk for k in alg if ku.get(k)}
            ka = train_step.get("fused_assembled_sh", {}).get("kernel_us_per_step", {})
            alg_a = {"assemble_forward": n_ray * 488 + n_bg * 476, "assemble_backward": P_GAUSS * 508}
            train_step["hbm_frac_assembled_sh"] = {k: alg_a[k] / (ka[k] * 1e-6) / 1e9 / pk for k in alg_a if ka.get(k)}
        except Exception as e:             # pragma: no cover  (never let the extra entry take the metric down)
            train_step = {"error": repr(e)[:300]}

    # ---------------- BASELINE config 2: the reference's own training resolution (README.md:65, -r 8) ----------------
    # 30k Gaussians, 504x378, SH 3 through the PUBLIC operator + the fused L1 + D-SSIM loss + autograd, a different camera
    # every step.  ~260 us of GPU work per step: the regime is host-bound, so what matters is whether host and GPU
    # overlap -- "fused" (default) blocks once per forward on R like the reference does, "async" never does.
    config2 = None
    if world == 1 and args.config == 3 and not args.no_config2:
        try:
            from scgaussian_b200 import synthetic as O2
            P2, W2, H2 = 30_000, 504, 378
            sc2 = O2.synth_scene(P2, W2, H2, sh_degree=3, scale_median=0.03, seed=0)
            lv = {k: sc2[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
            cams2 = [O2.make_camera(W2, H2, w2c=O2.yaw_w2c(camera_yaw(k))) for k in range(N_CAMERAS)]
            bg2 = torch.zeros(3, device=dev)
            from scgaussian_b200 import GaussianRasterizationSettings as GRS
            st2 = [GRS(image_height=H2, image_width=W2, tanfovx=c["tanfovx"], tanfovy=c["tanfovy"], bg=bg2, scale_modifier=1.0,
                       viewmatrix=c["viewmatrix"].to(dev), projmatrix=c["projmatrix"].to(dev), sh_degree=3,
                       campos=c["campos"].to(dev), prefiltered=False, debug=False) for c in cams2]
            gt2 = torch.rand(3, H2, W2, device=dev)
            m2 = torch.zeros(P2, 3, device=dev, requires_grad=True)

            def small_step(i):
                color, radii, depth, alpha = GaussianRasterizer(st2[i % N_CAMERAS])(
                    means3D=lv["means3D"], means2D=m2, opacities=lv["opacities"], shs=lv["shs"], scales=lv["scales"],
                    rotations=lv["rotations"])
                photometric_loss(color, gt2, 0.2).backward()
                for v in lv.values():
                    v.grad = None
                m2.grad = None

            config2 = {"workload": "config2: synthetic 30k Gaussians, 504x378, SH deg 3, public operator + fused L1/D-SSIM loss + autograd",
                       "steps": 300}
            saved_mode = R._BINNING_MODE
            for mode in ("fused", "async"):
                R._BINNING_MODE = mode
                for i in range(30):
                    small_step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for i in range(300):
                    small_step(i)
                e1.record()
                torch.cuda.synchronize()
                config2[mode] = {"us_per_step": max(e0.elapsed_time(e1) * 1e3, (time.perf_counter() - t0) * 1e6) / 300,
                                 "steps_per_s": 300.0 / max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)}
            config2["dropped_views_async"] = int(R.dropped_views)
            R._BINNING_MODE = saved_mode
            # the same step captured once and replayed as ONE CUDA graph launch (scgaussian_b200/graphs.py): the camera of
            # the step is copied into the captured tensors before every replay
            try:
                from scgaussian_b200.graphs import GraphedStep
                cam_t = {k: cams2[0][k].to(dev).clone() for k in ("viewmatrix", "projmatrix", "campos")}
                cam_src = [{k: c[k].to(dev) for k in cam_t} for c in cams2]
                sg = st2[0]._replace(viewmatrix=cam_t["viewmatrix"], projmatrix=cam_t["projmatrix"], campos=cam_t["campos"])

                def graph_body():
                    color, radii, depth, alpha = GaussianRasterizer(sg)(
                        means3D=lv["means3D"], means2D=m2, opacities=lv["opacities"], shs=lv["shs"], scales=lv["scales"],
                        rotations=lv["rotations"])
                    photometric_loss(color, gt2, 0.2).backward()

                for v in lv.values():
                    v.grad = None
                m2.grad = None
                gstep = GraphedStep(graph_body, warmup=3, device=dev)

                def graphed(i):
                    for k in cam_t:
                        cam_t[k].copy_(cam_src[i % N_CAMERAS][k], non_blocking=True)
                    for v in lv.values():
                        v.grad.zero_()
                    gstep.replay()

                for i in range(30):
                    graphed(i)
                torch.cuda.synchronize()
                # same numbers as the eager path?  (camera 5: gradients of the graph replay against an eager step)
                graphed(5)
                g_graph = {k: v.grad.clone() for k, v in lv.items()}
                for v in lv.values():
                    v.grad = None
                small_step(5)      # leaves grads None afterwards: recompute eagerly, keeping them
                color, radii, depth, alpha = GaussianRasterizer(st2[5])(means3D=lv["means3D"], means2D=m2, opacities=lv["opacities"],
                                                                      shs=lv["shs"], scales=lv["scales"], rotations=lv["rotations"])
                photometric_loss(color, gt2, 0.2).backward()
                worst = max(float((g_graph[k] - v.grad).abs().max()) / (float(v.grad.abs().max()) + 1e-30) for k, v in lv.items())
                for k, v in lv.items():
                    v.grad = g_graph[k]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for i in range(300):
                    graphed(i)
                e1.record()
                torch.cuda.synchronize()
                config2["graph"] = {"us_per_step": max(e0.elapsed_time(e1) * 1e3, (time.perf_counter() - t0) * 1e6) / 300,
                                    "steps_per_s": 300.0 / max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0),
                                    "overflowed": gstep.overflowed(), "max_rel_grad_diff_vs_eager": worst,
                                    "what": "forward + loss + backward captured once (async binning), replayed with one cudaGraphLaunch per step; "
                                            "the camera is copied into the captured tensors and the gradients are zeroed before every replay"}
            except Exception as e:         # pragma: no cover
                config2["graph"] = {"error": repr(e)[:300]}
            del lv, m2, gt2
        except Exception as e:             # pragma: no cover
            config2 = {"error": repr(e)[:300]}

    # ---------------- the naive-CUDA stand-in on the same inputs, same GPU (NOT the reference) ----------------
    standin = None
    if world == 1 and not args.no_standin and args.config == 3:
        try:
            from baseline.standin import standin as SB
            torch.cuda.empty_cache()
            standin = SB.measure(wl.scene_cpu, wl.cams_cpu, [g.cpu() for g in wl.grads], WIDTH, HEIGHT, SH_DEG, dev,
                                 steps=max(4, K // 3), warm=2)
            standin["speedup_of_this_repo"] = standin["ms_per_step"] / ms_step
        except Exception as e:             # pragma: no cover
            standin = {"error": repr(e)[:300]}

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = peaks()
    N = WIDTH * HEIGHT
    Tn = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
    algo_bytes = {   # SURVEY.md section 8d, per launch (one view), at the mean R of the profiled cameras
        "preprocess_forward": 311 * P_GAUSS,
        "render_forward": 44 * R_prof + 8 * Tn + 24 * N,
        "render_backward": 84 * R_prof + 8 * Tn + 28 * N,
        "preprocess_backward": 579 * P_GAUSS,
    }
    roof = None
    roof_issue = None
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        share = kernels[dom]["ms_per_step"] / max(sum(v["ms_per_step"] for v in kernels.values()), 1e-9)
        if dom in algo_bytes:
            ach = algo_bytes[dom] / (kernels[dom]["ms_per_launch"] * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(dom)
                except Exception:
                    traffic = None
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": algo_bytes[dom],
                    "share_of_step": share,
                    "note": "render kernels are SM-issue bound, not HBM bound (SURVEY 8d; see roofline_issue); HBM fraction reported as BASELINE.json requires"}
        for k, v in kernels.items():
            if k in algo_bytes:
                v["hbm_gbs_algorithmic"] = algo_bytes[k] / (v["ms_per_launch"] * 1e-3) / 1e9
                v["hbm_frac"] = v["hbm_gbs_algorithmic"] / peak
        if "preprocess_backward" in kernels:
            # SURVEY's 579 B/Gaussian assumes every Gaussian is processed; the kernel skips the ones render-backward left
            # without gradient (they are read as 48 B of accumulator and written as 232 B of zeros): the honest byte
            # count is the live-weighted one, and the nominal fraction (which can exceed 1) is labelled as such
            v = kernels["preprocess_backward"]
            v["hbm_frac_nominal_579B_per_gaussian"] = v.pop("hbm_frac")
            live_bytes = P_GAUSS * (live_frac * 579 + (1.0 - live_frac) * (48 + 232 + 4))
            v["live_fraction"] = live_frac
            v["hbm_gbs_algorithmic"] = live_bytes / (v["ms_per_launch"] * 1e-3) / 1e9
            v["hbm_frac"] = v["hbm_gbs_algorithmic"] / peak
        # ---- the bound that actually applies to the render kernels: SM issue slots ----
        # warp instructions per launch come from an ncu capture of this same command (profiles/r02_issue.json,
        # smsp__inst_executed.sum, averaged over the 8 cameras); durations are the live CUDA-event ones above
        f_clk = (clocks or {}).get("sm_mhz") or 1965.0
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sm_count * 4 * f_clk * 1e6            # warp instructions / s: one per scheduler per cycle
        lane_peak = sm_count * 128 * f_clk * 1e6           # 148 SMs x 128 lanes x f_clk
        counts = {}
        ip = os.path.join(ROOT, "profiles", "r02_issue.json")
        if os.path.exists(ip):
            try:
                counts = json.load(open(ip))
            except Exception:
                counts = {}
        roof_issue = {"peak_warp_instr_per_s": issue_peak, "sm_clock_mhz_used": f_clk, "sm_count": sm_count,
                      "source_of_instruction_counts": "profiles/r02_issue.json (ncu smsp__inst_executed.sum per launch)" if counts else None}
        for kname in ("render_forward", "render_backward"):
            if kname not in kernels:
                continue
            t_s = kernels[kname]["ms_per_launch"] * 1e-3
            pair_evals = 256.0 * R_prof                    # every pixel of a tile against every entry of its list (A.8)
            ent = {"ms_per_launch": kernels[kname]["ms_per_launch"], "pair_evals_per_s": pair_evals / t_s,
                   "lane_cycles_per_pair_eval": lane_peak * t_s / pair_evals}
            wi = (counts.get(kname) or {}).get("warp_instr_per_launch")
            if wi:
                ent.update({"warp_instr_per_launch": wi, "warp_instr_per_s": wi / t_s, "frac_of_issue_peak": wi / t_s / issue_peak,
                            "warp_instr_per_list_entry": wi / R_prof})
            roof_issue[kname] = ent

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import c_oracle
        c_oracle.build()
        times, cores = cpu_oracle_views_per_s(3, cfg)
        cpu = {"value": 1.0 / (sum(times[1:]) / len(times[1:])), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 full views fwd+bwd ({P_GAUSS} Gaussians, {WIDTH}x{HEIGHT}) after 1 warm-up, scalar C oracle, OpenMP over {cores} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "P": P_GAUSS, "width": WIDTH, "height": HEIGHT, "sh_degree": SH_DEG,
                   "cameras": f"{N_CAMERAS} yawed cameras ((k - 3.5) * 2 deg), rank r renders camera (r + step) % {N_CAMERAS}: a different view every step",
                   "num_rendered_R_mean": R_mean, "num_rendered_R_min": min(R_list), "num_rendered_R_max": max(R_list),
                   "num_rendered_distinct_over_steps": len(set(R_list)), "need_capacity_hits": int(need_capacity_hits),
                   "views_per_step": world, "parallelism": f"view-sharded dp{world}",
                   "collective": f"1 all-reduce of the flat gradient buffer per step ({wl.flat.collective if wl.flat is not None else 'nvls/nccl'})" if world > 1 else "none (1 GPU)",
                   "grad_allreduce_bytes": (P_GAUSS * 61 * 4) if world > 1 else 0,
                   "l2": "no explicit flush: one step streams >1 GB (inputs 236 MB + gradients 244 MB + scratch) through the 126 MB L2, and the camera changes every step",
                   "launch": "kernels of a stage chained by programmatic dependent launch (SCGR_PDL=%s)" % os.environ.get("SCGR_PDL", "1"),
                   "loss_kernels": "streaming column strips" if os.environ.get("SCGR_LOSS_VARIANT", "1") == "1" else "32x32 tiles"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "roofline_issue": roof_issue,
        "kernels": kernels, "batch8": batch8, "config4": config4, "allreduce_check": allreduce_check,
        "config2": config2, "gpu_standin_baseline": standin, "cpu_baseline": cpu, "train_step": train_step,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
