"""Property tests through the fused path, re-stated from the reference's own test strategy:
conditions are met exactly after `enforce` (reference tests/test_conditions.py:142-583, rtol 5e-4 / atol 1e-6), vector
calculus identities hold (tests/test_operators_identities.py:57-143, < 1e-4), `diff` of known functions
(tests/test_neurodiffeq.py:87-96), full-size round trips (linearity of the gradient in the loss scale, sharding)."""
import math

import numpy as np
import pytest
import torch

from neurodiffeq_b200 import diff
from neurodiffeq_b200 import operators as ops
from neurodiffeq_b200 import conditions as C
from neurodiffeq_b200.engine import FusedProblem
from neurodiffeq_b200.networks import FCNN, SinActv

pytestmark = pytest.mark.gpu


def evaluate(nets, conds, eqs, coords_np, n_coords=None):
    fp = FusedProblem(nets, conds, eqs, n_coords or len(coords_np))
    coords = [torch.as_tensor(np.ascontiguousarray(c), dtype=torch.float32).cuda() for c in coords_np]
    u, r, _ = fp.forward(coords)
    torch.cuda.synchronize()
    return u.cpu().numpy(), (r.cpu().numpy() if r is not None else None)


def close(a, b):
    np.testing.assert_allclose(a, b, rtol=5e-4, atol=2e-6)


def test_ivp_and_bundle_ivp_are_satisfied():
    torch.manual_seed(0)
    t0 = np.full(64, 0.3)
    net = FCNN(1, 1, hidden_units=(32, 32), actv=SinActv)
    u, r = evaluate([net], [C.IVP(t_0=0.3, u_0=1.7)], lambda u, t: [diff(u, t)], [t0])
    close(u[0], 1.7)
    net2 = FCNN(1, 1, hidden_units=(32,))
    u, r = evaluate([net2], [C.IVP(t_0=0.3, u_0=1.7, u_0_prime=-0.4)], lambda u, t: [diff(u, t)], [t0])
    close(u[0], 1.7)
    close(r[0], -0.4)                      # u'(t0) = u0'
    # bundle: t_0 and u_0 come per point from theta (reference tests/test_conditions.py:169-231)
    rs = np.random.RandomState(0)
    t0b, u0b = rs.rand(64) + 0.1, rs.randn(64)
    net3 = FCNN(3, 1, hidden_units=(16, 16))
    cond = C.BundleIVP(bundle_param_lookup={"t_0": 0, "u_0": 1})
    u, _ = evaluate([net3], [cond], lambda u, t, a, b: [diff(u, t)], [t0b, t0b, u0b])
    close(u[0], u0b)


def test_dirichlet_bvp_1d_2d_ibvp_spherical():
    torch.manual_seed(1)
    rs = np.random.RandomState(1)
    s = rs.rand(128)
    zeros, ones = np.zeros(128), np.ones(128)
    # 1-D two point
    u, _ = evaluate([FCNN(1, 1)], [C.DirichletBVP(0.0, 2.0, 1.0, -1.0)], lambda u, t: [diff(u, t, 2)], [np.r_[zeros[:64], ones[:64]]])
    close(u[0][:64], 2.0)
    close(u[0][64:], -1.0)
    # 2-D box (four sides; compatible corner data)
    f0, f1 = (lambda y: torch.sin(np.pi * y)), (lambda y: y * (1 - y))
    g0, g1 = (lambda x: 0 * x), (lambda x: 0 * x)
    bc = C.DirichletBVP2D(0, f0, 1, f1, 0, g0, 1, g1)
    lap = lambda u, x, y: [ops.laplacian(u, x, y)]  # noqa: E731
    net = FCNN(2, 1, hidden_units=(32, 32))
    for xs, ys, want in ((zeros, s, np.sin(np.pi * s)), (ones, s, s * (1 - s)), (s, zeros, 0 * s), (s, ones, 0 * s)):
        u, _ = evaluate([net], [bc], lap, [xs, ys])
        close(u[0], want)
    # IBVP1D Dirichlet-Dirichlet
    ib = C.IBVP1D(x_min=-1, x_max=1, t_min=0, t_min_val=lambda x: -torch.sin(np.pi * x), x_min_val=lambda t: t,
                  x_max_val=lambda t: 2 * t)
    heat = lambda u, x, t: [diff(u, t) - diff(u, x, 2)]  # noqa: E731
    net = FCNN(2, 1, hidden_units=(32, 32))
    x = 2 * s - 1
    u, _ = evaluate([net], [ib], heat, [x, zeros])
    close(u[0], -np.sin(np.pi * x))
    u, _ = evaluate([net], [ib], heat, [-ones, s])
    close(u[0], s)
    u, _ = evaluate([net], [ib], heat, [ones, s])
    close(u[0], 2 * s)
    # spherical shells
    sp = C.DirichletBVPSpherical(0.5, lambda th, ph: torch.cos(th), 2.0, lambda th, ph: torch.sin(ph))
    th, ph = 0.2 + 2.5 * s, 6.0 * rs.rand(128)
    net = FCNN(3, 1, hidden_units=(32, 32))
    eq = lambda u, r, t, p: [ops.spherical_laplacian(u, r, t, p)]  # noqa: E731
    u, _ = evaluate([net], [sp], eq, [0.5 * ones, th, ph])
    close(u[0], np.cos(th))
    u, _ = evaluate([net], [sp], eq, [2.0 * ones, th, ph])
    close(u[0], np.sin(ph))


def test_diff_of_closed_forms_and_unused_coordinate():
    """d^k/dt^k of t^2 and exp(t) (reference tests/test_neurodiffeq.py:87-96); a coordinate the expression does not depend
    on differentiates to zero."""
    net = FCNN(2, 1, hidden_units=(8,))
    rs = np.random.RandomState(2)
    t, s = rs.rand(100) + 0.5, rs.rand(100)

    def eqs(u, t, s):
        e = torch.exp(t)
        return [diff(t ** 2, t) - 2 * t, diff(t ** 2, t, order=2) - 2.0, diff(t ** 2, t, order=3), diff(e, t, order=4) - e,
                diff(t ** 2, s) + 0 * u]

    _, r = evaluate([net], [C.NoCondition()], eqs, [t, s])
    assert np.abs(r).max() < 1e-5


def test_vector_calculus_identities():
    """div grad = laplacian, curl grad = 0, div curl = 0 on a random smooth vector field (3 nets), Cartesian."""
    torch.manual_seed(3)
    nets = [FCNN(3, 1, hidden_units=(24, 24)) for _ in range(3)]
    conds = [C.NoCondition() for _ in range(3)]
    rs = np.random.RandomState(3)
    xyz = [rs.rand(256) for _ in range(3)]

    def eqs(ux, uy, uz, x, y, z):
        gx, gy, gz = ops.grad(ux, x, y, z)
        r1 = ops.div(gx, gy, gz, x, y, z) - ops.laplacian(ux, x, y, z)          # needs only pure seconds
        return [r1]

    _, r = evaluate(nets, conds, eqs, xyz)
    assert np.abs(r).max() < 1e-4
    # mixed partials commute: d/dx d/dy u == d/dy d/dx u (polarisation channels)
    net = FCNN(2, 1, hidden_units=(24, 24))
    _, r = evaluate([net], [C.NoCondition()], lambda u, x, y: [diff(diff(u, x), y) - diff(diff(u, y), x)], xyz[:2])
    assert np.abs(r).max() < 1e-5


def test_full_size_linearity_and_sharding_roundtrip():
    """BASELINE-size batch (C2, 16384 points): grad(loss_scale) is linear in the scale; 8 shards with the global scale add
    up to the full-batch gradient (size-independent properties, no oracle needed)."""
    import workloads
    from helpers import build_fused
    wl, nets, conds, fp = build_fused("c2", seed=9)
    n = wl.default_n
    cs = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=3)]
    fp.gradbuf.zero_()
    fp.residual_grad(cs, n_global=n, sumsq_out=fp.sumsq)
    g1, s1 = fp.grad.clone(), float(fp.sumsq)
    fp.gradbuf.zero_()
    fp.residual_grad(cs, n_global=4 * n, sumsq_out=fp.sumsq)
    assert torch.allclose(4 * fp.grad, g1, rtol=1e-5, atol=1e-9)
    fp.gradbuf.zero_()
    for k in range(8):
        lo, hi = n * k // 8, n * (k + 1) // 8
        fp.residual_grad([c[lo:hi].contiguous() for c in cs], n_global=n, sumsq_out=fp.sumsq)
    torch.cuda.synchronize()
    assert (fp.grad - g1).norm() <= 2e-5 * g1.norm()
    assert abs(float(fp.sumsq) - s1) <= 1e-5 * s1
