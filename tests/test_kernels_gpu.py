"""Parity of the CUDA path (through the C ABI, libpinnjet.so) against
 (a) golden vectors produced by the unmodified reference (tests/golden/*.npz), and
 (b) the CPU oracle on fresh seeded inputs at sizes that are not multiples of the tile,
for all five BASELINE.json workloads.  Tolerances: helpers.TOL_* (fp32 kernels vs fp64 reference)."""
import numpy as np
import pytest
import torch

import workloads
from conftest import load_golden
from helpers import build_fused, oracle_eval, get_params, assert_parity

pytestmark = pytest.mark.gpu


def run_fused(fp, coords_np, n_global=None):
    coords = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in coords_np]
    u, r, sumsq = fp.forward(coords, want_sumsq=True)
    n = coords_np.shape[1]
    loss_eval = float(sumsq.item()) / (n * fp.n_eq)
    fp.grad.zero_()
    s2, r2 = fp.residual_grad(coords, n_global=n_global, want_residual=True)
    torch.cuda.synchronize()
    loss_train = float(s2.item()) / (n * fp.n_eq)
    return (u.cpu().numpy(), r.cpu().numpy(), loss_eval, r2.cpu().numpy(), loss_train, fp.grads_as_list())


@pytest.mark.parametrize("key", workloads.NAMES)
def test_matches_reference_golden(key):
    wl0 = workloads.build(__import__("helpers").product_namespace(), key)
    gold = load_golden(wl0.name)
    wl, nets, conds, fp = build_fused(key, params=gold["params"])
    u, r, loss_eval, r2, loss_train, grads = run_fused(fp, gold["coords"])
    assert_parity(u, r, loss_eval, grads, gold, label=f"{key} golden")
    assert_parity(None, r2, loss_train, None, gold, label=f"{key} golden(train fwd)")


@pytest.mark.parametrize("key,n", [("c1", 1024), ("c2", 5000), ("c3", 3001), ("c4", 4097), ("c5", 10007)])
def test_matches_oracle_ragged_sizes(key, n):
    wl, nets, conds, fp = build_fused(key, seed=7)
    params = get_params(nets)
    coords = workloads.sample_coords(wl, n, seed=99)
    ref = oracle_eval(key, params, coords)
    ref["residual32"] = oracle_eval(key, params, coords, dtype=torch.float32, backward=False)["residual"]
    u, r, loss_eval, r2, loss_train, grads = run_fused(fp, coords)
    assert_parity(u, r, loss_eval, grads, ref, label=f"{key} N={n}")
    assert_parity(None, r2, loss_train, None, ref, label=f"{key} N={n} (train fwd)")


def test_gradient_accumulates_and_linearity():
    """grad_theta is ACCUMULATED (reference solvers.py:360-362); two half batches with n_global = N sum to the full
    batch gradient (the multi-GPU sharding identity)."""
    wl, nets, conds, fp = build_fused("c2", seed=3)
    n = 4096
    coords_np = workloads.sample_coords(wl, n, seed=5)
    coords = [torch.from_numpy(c).cuda() for c in coords_np]
    fp.grad.zero_()
    fp.residual_grad(coords)
    g_full = fp.grad.clone()
    fp.grad.zero_()
    fp.residual_grad([c[: n // 2].contiguous() for c in coords], n_global=n)
    fp.residual_grad([c[n // 2:].contiguous() for c in coords], n_global=n)
    torch.cuda.synchronize()
    rel = (fp.grad - g_full).norm() / g_full.norm()
    assert rel < 1e-5, rel
    fp.residual_grad(coords)   # accumulate on top
    rel2 = (fp.grad - 2 * g_full).norm() / g_full.norm()
    assert rel2 < 1e-5, rel2


def test_single_point_and_tiny_batches():
    wl, nets, conds, fp = build_fused("c2", seed=11)
    params = get_params(nets)
    for n in (1, 2, 31, 33):
        coords = workloads.sample_coords(wl, n, seed=n)
        ref = oracle_eval("c2", params, coords)
        u, r, loss_eval, r2, loss_train, grads = run_fused(fp, coords)
        assert_parity(u, r, loss_eval, grads, ref, label=f"c2 N={n}")


def test_library_is_the_path():
    """The product must be running out of the in-tree shared library (no eager / CPU fallback)."""
    from neurodiffeq_b200 import engine
    import os
    assert os.path.exists(engine.library_path())
    with open("/proc/self/maps") as f:
        assert "libpinnjet.so" in f.read()


@pytest.mark.parametrize("key", ["c1", "c2", "c3"])
def test_pack_clears_the_gradient_buffer_and_the_loss_is_finalised_in_kernel(key):
    """K0 with zero_gradbuf (pj_pack_zero) packs like pj_pack and clears [grad | sum r^2]; the forward kernel's last warp
    folds the per-CTA partial sums itself (ticket re-armed for the next launch): repeated launches give identical sums."""
    wl, nets, conds, fp = build_fused(key, seed=3)
    n = 5000
    coords = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=5)]
    fp.pack()
    torch.cuda.synchronize()
    packed = fp.pack_buf.clone()
    fp.gradbuf.fill_(7.0)
    fp.pack_buf.zero_()                     # regions no kernel reads (padding, images of the other kernel family) stay zero
    fp.pack(zero_gradbuf=True)
    torch.cuda.synchronize()
    assert torch.equal(fp.pack_buf, packed)
    assert float(fp.gradbuf.abs().max()) == 0.0
    sums = []
    for _ in range(6):
        _, r, s = fp.forward(coords, want_u=False, want_residual=True, want_sumsq=True)
        sums.append(float(s.item()))
    ref = float((r.double() ** 2).sum())
    assert len(set(sums)) == 1 and abs(sums[0] - ref) <= 1e-5 * ref
    fp.gradbuf.zero_()
    fp.residual_grad(coords, sumsq_out=fp.sumsq)
    torch.cuda.synchronize()
    g1 = fp.gradbuf.clone()
    for _ in range(3):      # clears the stale content, never accumulates
        fp.gradbuf.fill_(3.0)
        fp.residual_grad(coords, sumsq_out=fp.sumsq, zero_gradbuf=True)
        torch.cuda.synchronize()
        assert torch.equal(fp.gradbuf, g1)
    fp.residual_grad(coords, sumsq_out=fp.sumsq)            # and without the flag it still accumulates
    torch.cuda.synchronize()
    np.testing.assert_allclose(fp.gradbuf.cpu().numpy(), 2.0 * g1.cpu().numpy(), rtol=1e-6, atol=1e-7)
