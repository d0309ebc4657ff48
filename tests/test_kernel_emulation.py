"""The kernels of scgaussian_b200/csrc/model.cu executed on the HOST, thread for thread (tests/emulation/: the .cu
source compiled with g++ through a small CUDA shim -- real threads per block, real barriers, shared memory as
statics), against the same oracle and reference-generated golden vectors as the GPU tests.

TEST INFRASTRUCTURE, CPU suite only: it catches indexing / bounds / table / staging mistakes before a GPU is
available.  It is NOT a CPU path of the product (nothing under scgaussian_b200/ can reach it) and it proves nothing
about the GPU build's numerics beyond what the arithmetic shares -- the `-m gpu` tests remain the parity gate."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as MO
from scgaussian_b200 import _lib as L
from tests import test_model as TM

GOLD = TM.GOLD


@pytest.fixture(scope="module")
def emu():
    from tests.emulation import build
    try:
        path = build.build()
    except Exception as e:      # pragma: no cover  (no g++ / no C++20 <barrier>)
        pytest.skip(f"host emulation library not buildable here: {e}")
    return C.CDLL(path)


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _model(ray, bg, K):
    tr = {k: torch.tensor(np.asarray(v), dtype=torch.float32) for k, v in ray.items()}
    tb = {k: torch.tensor(np.asarray(v), dtype=torch.float32) for k, v in bg.items()}

    def mk(d):
        if not d:
            return L.ScgrModelSet(0, *[None] * 9)
        return L.ScgrModelSet(int(d["scaling"].shape[0]), _p(d.get("xyz")), _p(d.get("rayo")), _p(d.get("rayd")),
                              _p(d.get("zval")), _p(d["scaling"]), _p(d["rotation"]), _p(d["opacity"]),
                              _p(d["features_dc"]), _p(d.get("features_rest")))
    return L.ScgrModel(K - 1, (L.ScgrModelSet * 2)(mk(tr), mk(tb))), (tr, tb)


# staged SH copy: K = 16, 4; flat copy: K = 9, 1; chunk boundaries of the staged copy: 63 / 64 / 65 / 128 Gaussians
@pytest.mark.parametrize("n_ray,n_bg,K", [(1, 1, 16), (0, 133, 16), (200, 0, 16), (321, 190, 16), (130, 65, 4),
                                          (64, 128, 16), (63, 0, 16), (77, 123, 9), (30, 21, 1)])
def test_assemble_kernels_on_host_match_oracle(emu, n_ray, n_bg, K):
    ray, bg = TM._random_sets(n_ray, n_bg, K, seed=n_ray + n_bg + K)
    if n_bg > 2:
        bg["opacity"][2] = 40.0
        bg["rotation"][1] = 0.0
    want, leaves = MO.assemble(ray, bg, torch.float32)
    model, keep = _model(ray, bg, K)
    P = n_ray + n_bg
    outs = [torch.full(s, float("nan")) for s in ((P, 3), (P, 3), (P, 4), (P, 1), (P, K, 3))]
    out = L.ScgrActivated(*[t.data_ptr() for t in outs])
    emu.emu_assemble_forward(C.byref(model), C.byref(out))
    for name, a, b in zip(TM.ACT, outs, want):
        assert not torch.isnan(a).any(), name
        assert TM._rel(a.numpy(), b.detach().numpy()) < TM.RTOL_ACT, name
    assert torch.equal(outs[4], want[4].detach())                       # the SH block is a pure copy
    ws = [torch.randn(a.shape, generator=torch.Generator().manual_seed(5)) for a in want]
    sum((a * w).sum() for a, w in zip(want, ws)).backward()
    gin = L.ScgrActivatedGrads(*[w.data_ptr() for w in ws])

    def grads(n, is_ray):
        nan = lambda *s: torch.full(s, float("nan"))     # noqa: E731
        return dict(xyz=None if is_ray else nan(n, 3), zval=nan(n, 1) if is_ray else None, scaling=nan(n, 3),
                    rotation=nan(n, 4), opacity=nan(n, 1), features_dc=nan(n, 1, 3), features_rest=nan(n, K - 1, 3))
    d_ray, d_bg = grads(n_ray, True), grads(n_bg, False)

    def gset(d, n):
        if not n:
            return L.ScgrModelSetGrads(*[None] * 7)
        return L.ScgrModelSetGrads(_p(d["xyz"]), _p(d["zval"]), _p(d["scaling"]), _p(d["rotation"]), _p(d["opacity"]),
                                   _p(d["features_dc"]), _p(d["features_rest"]))
    og = L.ScgrModelGrads((L.ScgrModelSetGrads * 2)(gset(d_ray, n_ray), gset(d_bg, n_bg)))
    emu.emu_assemble_backward(C.byref(model), C.byref(gin), C.byref(og))
    for prefix, d, n in (("ray_", d_ray, n_ray), ("bg_", d_bg, n_bg)):
        for k, t in d.items():
            if t is None or not n:
                continue
            assert not torch.isnan(t).any(), prefix + k                  # every output element written
            assert TM._rel(t.numpy(), leaves[prefix + k].grad.numpy()) < TM.RTOL_GRAD, prefix + k


def test_adam_kernel_on_host_replays_reference_golden(emu):
    g = np.load(GOLD)
    ps = {n: torch.tensor(g[f"adam_p0{n}"]) for n in TM.TRAINED}
    ms = {n: torch.zeros_like(ps[n]) for n in ps}
    vs = {n: torch.zeros_like(ps[n]) for n in ps}
    for it in (1, 2, 3):
        gr = {n: torch.tensor(g[f"adam_g{it}{n}"]) for n in ps}
        tab = (L.ScgrAdamGroup * 12)(*[L.ScgrAdamGroup(ps[n].data_ptr(), gr[n].data_ptr(), ms[n].data_ptr(), vs[n].data_ptr(),
                                                       ps[n].numel(), float(g[f"adam_lr{n}"][it - 1]), it) for n in TM.TRAINED])
        emu.emu_adam_step(tab, 12, C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        for n in ps:
            want = g[f"adam_p{it}{n}"]
            assert np.abs(ps[n].numpy() - want).max() <= 1e-6 * np.abs(want).max(), (n, it)
    for n in ps:
        assert TM._rel(ms[n].numpy(), g[f"adam_m3{n}"]) < 1e-6 and TM._rel(vs[n].numpy(), g[f"adam_v3{n}"]) < 1e-6, n


def test_adam_kernel_on_host_ragged_and_unaligned(emu):
    gen = torch.Generator().manual_seed(2)
    sizes = [(1,), (3,), (4096,), (4097,), (9001, 3), (8191,)]
    lrs = [0.5, 1e-3, 1.6e-4, 5.5e-2, 2e-3, 1.5e-3]
    inits = [torch.randn(*s, generator=gen) for s in sizes]
    backing = torch.zeros(9000)
    ours = [t.clone() for t in inits]
    backing[1:8192].copy_(inits[5])
    ours[5] = backing[1:8192]                       # 4 bytes off 16-byte alignment: the scalar path
    assert ours[5].data_ptr() % 16 == 4
    ref_p = [torch.nn.Parameter(t.clone()) for t in inits]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref_p, lrs)], lr=0.0, eps=1e-15)
    M, V = [torch.zeros_like(t) for t in inits], [torch.zeros_like(t) for t in inits]
    for it in range(1, 4):
        grs = [torch.randn(t.shape, generator=gen) * 10.0 ** ((k % 4) - 3) for k, t in enumerate(inits)]
        for p, gr in zip(ref_p, grs):
            p.grad = gr.clone()
        ref.step()
        tab = (L.ScgrAdamGroup * len(ours))(*[L.ScgrAdamGroup(ours[k].data_ptr(), grs[k].data_ptr(), M[k].data_ptr(),
                                                              V[k].data_ptr(), ours[k].numel(), lrs[k], it)
                                              for k in range(len(ours))])
        emu.emu_adam_step(tab, len(ours), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        for k in range(len(ours)):
            scale = float(ref_p[k].detach().abs().max()) + lrs[k]
            assert float((ours[k] - ref_p[k].detach()).abs().max()) <= 2e-6 * scale, (it, k)
    assert float(backing[0]) == 0.0 and float(backing[8192:].abs().max()) == 0.0      # nothing written outside the view


def test_densification_stats_kernel_on_host_matches_reference_golden(emu):
    g = np.load(GOLD)
    a, d, m = (torch.tensor(g[k]) for k in ("stats_accum0", "stats_denom0", "stats_maxr0"))
    gr, r = torch.tensor(g["stats_grad"]), torch.tensor(g["stats_radii"])
    emu.emu_densification_stats(_p(gr), None, _p(r), len(r), _p(a), _p(d), _p(m))
    assert np.array_equal(d.numpy(), g["stats_denom1"]) and np.array_equal(m.numpy(), g["stats_maxr1"])
    assert TM._rel(a.numpy(), g["stats_accum1"]) < 1e-6
    # explicit filter, no radii: max_radii2D untouched
    a, d = torch.tensor(g["stats_accum0"]), torch.tensor(g["stats_denom0"])
    f = r > 7
    emu.emu_densification_stats(_p(gr), _p(f), None, len(r), _p(a), _p(d), None)
    wa, wd, _ = MO.densification_stats(g["stats_grad"], f.numpy(), None, g["stats_accum0"], g["stats_denom0"], g["stats_maxr0"])
    assert np.array_equal(d.numpy(), wd) and TM._rel(a.numpy(), wa) < 1e-6


@pytest.mark.parametrize("n_in,keep", [(1, 1.0), (700, 0.5), (9001, 0.9), (500, 0.0), (3333, 0.01)])
def test_gather_rows_kernel_on_host_matches_torch_indexing(emu, n_in, keep):
    gen = torch.Generator().manual_seed(n_in)
    mask = torch.rand(n_in, generator=gen) < keep
    idx = mask.nonzero().squeeze(1)
    shapes = [(1,), (3,), (1, 3), (15, 3), (4,), (), (300,), (257,)]
    srcs = [torch.randn(n_in, *s, generator=gen) for s in shapes]
    dsts = [torch.full((len(idx), *s), float("nan")) for s in shapes]
    if len(idx):
        tab = (L.ScgrRowGather * len(srcs))(*[L.ScgrRowGather(a.data_ptr(), b.data_ptr(), a.numel() // n_in)
                                              for a, b in zip(srcs, dsts)])
        emu.emu_gather_rows(tab, len(srcs), _p(idx), C.c_int64(len(idx)))
    for a, b in zip(srcs, dsts):
        assert torch.equal(b, a[mask])


def test_copy_segments_kernel_on_host(emu):
    gen = torch.Generator().manual_seed(0)
    big = torch.full((40000,), float("nan"))
    segs, off = [], 1
    for k, n in enumerate([1, 3, 4095, 4096, 4097, 12289, 0, 8192]):
        src = torch.randn(n, generator=gen) if k % 3 != 2 else None            # every third segment is a zero fill
        if k % 2 == 0:
            off += (4 - off % 4) % 4                                             # 16-byte aligned destination
        segs.append((off, n, src))
        off += n + 3
    tab = (L.ScgrSegmentCopy * len(segs))(*[L.ScgrSegmentCopy(big[o:o + n].data_ptr() if n else None,
                                                               None if s is None else s.data_ptr(), n) for o, n, s in segs])
    emu.emu_copy_segments(tab, len(segs))
    covered = torch.zeros(40000, dtype=torch.bool)
    for o, n, s in segs:
        assert torch.equal(big[o:o + n], s if s is not None else torch.zeros(n)), (o, n)
        covered[o:o + n] = True
    assert torch.isnan(big[~covered]).all()                                      # nothing written outside the segments
    assert {o % 4 for o, n, s in segs if n} > {0}                                # both the float4 and the scalar path ran
