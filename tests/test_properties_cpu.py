"""Property tests of the traced problem definitions on the CPU (float64 stand-in engine, tests/cpu_engine.py), re-stated from
the reference's test strategy: every condition is met after `enforce` for ANY weights (reference
tests/test_conditions.py:142-583), including the Neumann ends that evaluate the network at a boundary abscissa; vector
calculus identities (tests/test_operators_identities.py:57-143); `diff` of closed forms (tests/test_neurodiffeq.py:87-96).
The GPU suite repeats the core of this through the CUDA kernels at fp32 tolerances (tests/test_properties_gpu.py)."""
import numpy as np
import torch

from cpu_engine import CpuFusedProblem
from neurodiffeq_b200 import diff
from neurodiffeq_b200 import operators as ops
from neurodiffeq_b200 import conditions as C
from neurodiffeq_b200.networks import FCNN, SinActv


def evaluate(nets, conds, eqs, coords_np):
    fp = CpuFusedProblem(nets, conds, eqs, len(coords_np))
    u, r, _ = fp.forward([torch.as_tensor(np.ascontiguousarray(c), dtype=torch.float64) for c in coords_np])
    return u.numpy(), (r.numpy() if r is not None else None)


def close(a, b, tol=2e-6):   # boundary data are lowered as float32 immediates: 1e-7-level agreement
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


VALUE_AND_SLOPE_1D = lambda u, x: [u, diff(u, x)]                # noqa: E731  "residuals" that expose u and u'
VALUE_AND_SLOPE_XT = lambda u, x, t: [u, diff(u, x)]             # noqa: E731


def test_initial_value_conditions():
    torch.manual_seed(0)
    t0 = np.full(16, 0.3)
    u, r = evaluate([FCNN(1, 1, hidden_units=(16, 16), actv=SinActv)], [C.IVP(t_0=0.3, u_0=1.7)], VALUE_AND_SLOPE_1D, [t0])
    close(u[0], 1.7)
    u, r = evaluate([FCNN(1, 1, hidden_units=(16,))], [C.IVP(t_0=0.3, u_0=1.7, u_0_prime=-0.4)], VALUE_AND_SLOPE_1D, [t0])
    close(u[0], 1.7)
    close(r[1], -0.4)
    rs = np.random.RandomState(0)
    t0b, u0b, v0b = rs.rand(16) + 0.1, rs.randn(16), rs.randn(16)
    cond = C.BundleIVP(bundle_param_lookup={"t_0": 0, "u_0": 1, "u_0_prime": 2})
    u, r = evaluate([FCNN(4, 1, hidden_units=(16, 16))], [cond], lambda u, t, a, b, c: [diff(u, t)], [t0b, t0b, u0b, v0b])
    close(u[0], u0b)
    close(r[0], v0b)


def test_two_point_conditions_with_every_dirichlet_neumann_combination():
    torch.manual_seed(1)
    ends = [np.zeros(8), np.ones(8)]
    x = np.r_[ends[0], ends[1]]
    lo, hi = slice(0, 8), slice(8, 16)
    net = lambda: FCNN(1, 1, hidden_units=(16, 16))              # noqa: E731
    u, _ = evaluate([net()], [C.DirichletBVP(0.0, 2.0, 1.0, -1.0)], VALUE_AND_SLOPE_1D, [x])
    close(u[0][lo], 2.0), close(u[0][hi], -1.0)
    rs = np.random.RandomState(1)
    a, b = rs.randn(16), rs.randn(16)
    bundle = C.BundleDirichletBVP(t_0=0.0, t_1=1.0, bundle_param_lookup={"u_0": 0, "u_1": 1})
    u, _ = evaluate([FCNN(3, 1, hidden_units=(16,))], [bundle], lambda u, t, p, q: [diff(u, t)], [x, a, b])
    close(u[0][lo], a[lo]), close(u[0][hi], b[hi])
    u, r = evaluate([net()], [C.DoubleEndedBVP1D(0.0, 1.0, x_min_val=2.0, x_max_val=-1.0)], VALUE_AND_SLOPE_1D, [x])
    close(r[0][lo], 2.0), close(r[0][hi], -1.0)
    u, r = evaluate([net()], [C.DoubleEndedBVP1D(0.0, 1.0, x_min_val=2.0, x_max_prime=0.7)], VALUE_AND_SLOPE_1D, [x])
    close(r[0][lo], 2.0), close(r[1][hi], 0.7)                   # u(x0), u'(x1)
    u, r = evaluate([net()], [C.DoubleEndedBVP1D(0.0, 1.0, x_min_prime=-0.3, x_max_val=1.5)], VALUE_AND_SLOPE_1D, [x])
    close(r[1][lo], -0.3), close(r[0][hi], 1.5)                  # u'(x0), u(x1)
    u, r = evaluate([net()], [C.DoubleEndedBVP1D(0.0, 1.0, x_min_prime=-0.3, x_max_prime=0.7)], VALUE_AND_SLOPE_1D, [x])
    close(r[1][lo], -0.3), close(r[1][hi], 0.7)


def test_ibvp1d_with_every_dirichlet_neumann_combination():
    torch.manual_seed(2)
    rs = np.random.RandomState(2)
    s = rs.rand(12)
    zeros, ones = np.zeros(12), np.ones(12)
    u0 = lambda x: torch.sin(0.5 * np.pi * x)                    # noqa: E731
    g, h = (lambda t: 0.2 * torch.sin(t)), (lambda t: 1.0 + 0.1 * t)   # Dirichlet data, compatible with u0 at t = 0
    p, q = (lambda t: 0.5 * np.pi + 0.3 * t), (lambda t: 0.1 * t)      # Neumann data, compatible with u0' at t = 0
    net = lambda: FCNN(2, 1, hidden_units=(16, 16))              # noqa: E731
    cases = {
        "dd": dict(x_min_val=g, x_max_val=h), "dn": dict(x_min_val=g, x_max_prime=q),
        "nd": dict(x_min_prime=p, x_max_val=h), "nn": dict(x_min_prime=p, x_max_prime=q),
    }
    for kind, kw in cases.items():
        cond = C.IBVP1D(x_min=0.0, x_max=1.0, t_min=0.0, t_min_val=u0, **kw)
        u, r = evaluate([net()], [cond], VALUE_AND_SLOPE_XT, [s, zeros])          # initial line
        close(r[0], np.sin(0.5 * np.pi * s))
        u, r = evaluate([net()], [cond], VALUE_AND_SLOPE_XT, [zeros, s])          # left end
        close(r[0] if kind[0] == "d" else r[1], 0.2 * np.sin(s) if kind[0] == "d" else 0.5 * np.pi + 0.3 * s)
        u, r = evaluate([net()], [cond], VALUE_AND_SLOPE_XT, [ones, s])           # right end
        close(r[0] if kind[1] == "d" else r[1], 1.0 + 0.1 * s if kind[1] == "d" else 0.1 * s)


def test_heat_equation_residual_with_neumann_data_on_both_ends_matches_autograd():
    """IBVP1D with two Neumann ends inside a PDE: mixed derivatives d2/dt dx0 and d2/dt dx1 of the two boundary instances
    ride on ONE shared polarisation direction; the residual equals the oracle's autograd residual."""
    from oracle import reference_port as oracle
    torch.manual_seed(6)
    net = FCNN(2, 1, hidden_units=(12, 12)).double()
    kw = dict(x_min=0.0, x_max=1.0, t_min=0.0, t_min_val=lambda x: torch.sin(0.5 * np.pi * x),
              x_min_prime=lambda t: 0.5 * np.pi + 0.1 * t, x_max_prime=lambda t: 0.3 * torch.sin(t))
    heat = lambda u, x, t: [diff(u, t) - 0.3 * diff(u, x, order=2)]          # noqa: E731
    rs = np.random.RandomState(6)
    xs, ts = rs.rand(30), rs.rand(30)
    _, r = evaluate([net], [C.IBVP1D(**kw)], heat, [xs, ts])
    cols = [torch.tensor(v).reshape(-1, 1).requires_grad_(True) for v in (xs, ts)]
    u = oracle.IBVP1D(**kw).enforce(net, *cols)
    ref = oracle.diff(u, cols[1]) - 0.3 * oracle.diff(u, cols[0], order=2)
    close(r[0], ref.detach().numpy()[:, 0], tol=1e-6)


def test_box_and_spherical_dirichlet_conditions():
    torch.manual_seed(3)
    rs = np.random.RandomState(3)
    s = rs.rand(10)
    zeros, ones = np.zeros(10), np.ones(10)
    f0, f1 = (lambda y: torch.sin(np.pi * y)), (lambda y: y * (1 - y))
    bc = C.DirichletBVP2D(0, f0, 1, f1, 0, lambda x: 0 * x, 1, lambda x: 0 * x)
    lap = lambda u, x, y: [ops.laplacian(u, x, y)]               # noqa: E731
    net = FCNN(2, 1, hidden_units=(16, 16))
    for xs, ys, want in ((zeros, s, np.sin(np.pi * s)), (ones, s, s * (1 - s)), (s, zeros, 0 * s), (s, ones, 0 * s)):
        close(evaluate([net], [bc], lap, [xs, ys])[0][0], want)
    th, ph = 0.2 + 2.5 * s, 6.0 * rs.rand(10)
    eq = lambda u, r, t, p: [ops.spherical_laplacian(u, r, t, p)]  # noqa: E731
    net3 = FCNN(3, 1, hidden_units=(16, 16))
    shell = C.DirichletBVPSpherical(0.5, lambda t, p: torch.cos(t), 2.0, lambda t, p: torch.sin(p))
    close(evaluate([net3], [shell], eq, [0.5 * ones, th, ph])[0][0], np.cos(th))
    close(evaluate([net3], [shell], eq, [2.0 * ones, th, ph])[0][0], np.sin(ph))
    one_sided = C.DirichletBVPSpherical(0.5, lambda t, p: torch.cos(t))
    close(evaluate([net3], [one_sided], eq, [0.5 * ones, th, ph])[0][0], np.cos(th))
    inf = C.InfDirichletBVPSpherical(0.5, lambda t, p: torch.cos(t), lambda t, p: torch.sin(p), order=2)
    close(evaluate([net3], [inf], eq, [0.5 * ones, th, ph])[0][0], np.cos(th))
    close(evaluate([net3], [inf], eq, [40.0 * ones, th, ph])[0][0], np.sin(ph))     # r -> infinity


def test_diff_of_closed_forms_and_unused_coordinate():
    net = FCNN(2, 1, hidden_units=(8,))
    rs = np.random.RandomState(4)
    t, s = rs.rand(20) + 0.5, rs.rand(20)

    def eqs(u, t, s):
        e = torch.exp(t)
        return [diff(t ** 2, t) - 2 * t, diff(t ** 2, t, order=2) - 2.0, diff(t ** 2, t, order=3), diff(e, t, order=4) - e,
                diff(t ** 2, s) + 0 * u, diff(torch.sin(t * s), t, order=2) + s * s * torch.sin(t * s)]

    _, r = evaluate([net], [C.NoCondition()], eqs, [t, s])
    assert np.abs(r).max() < 1e-6


def test_vector_calculus_identities():
    torch.manual_seed(5)
    nets = [FCNN(3, 1, hidden_units=(12, 12)) for _ in range(3)]
    conds = [C.NoCondition() for _ in range(3)]
    rs = np.random.RandomState(5)
    xyz = [rs.rand(24) for _ in range(3)]

    def div_grad(ux, uy, uz, x, y, z):
        gx, gy, gz = ops.grad(ux, x, y, z)
        return [ops.div(gx, gy, gz, x, y, z) - ops.laplacian(ux, x, y, z)]

    _, r = evaluate(nets, conds, div_grad, xyz)
    assert np.abs(r).max() < 1e-9
    net = FCNN(2, 1, hidden_units=(12, 12))
    _, r = evaluate([net], [C.NoCondition()], lambda u, x, y: [diff(diff(u, x), y) - diff(diff(u, y), x)], xyz[:2])
    assert np.abs(r).max() < 1e-12

    # spherical: the Laplacian of a radial function f(r) is f'' + 2 f'/r, whatever the angles
    def radial(u, r, th, ph):
        f = torch.exp(-r * r)
        return [ops.spherical_laplacian(f, r, th, ph) - (diff(f, r, order=2) + 2 * diff(f, r) / r) + 0 * u]

    _, r = evaluate([FCNN(3, 1, hidden_units=(8,))], [C.NoCondition()], radial, [xyz[0] + 0.5, xyz[1] + 0.3, xyz[2]])
    assert np.abs(r).max() < 1e-6


def test_resnet_with_a_neumann_boundary_instance_matches_autograd():
    """A Resnet evaluated at the sample points AND at a boundary abscissa: two instances share the body's weights and the
    shortcut matrix; loss and every parameter gradient equal autograd's (oracle)."""
    from oracle import reference_port as oracle
    from neurodiffeq_b200.networks import Resnet
    torch.manual_seed(8)
    net = Resnet(2, 1, hidden_units=(10, 10)).double()
    kw = dict(x_min=0.0, x_max=1.0, t_min=0.0, t_min_val=lambda x: torch.sin(0.5 * np.pi * x),
              x_min_val=lambda t: 0.2 * torch.sin(t), x_max_prime=lambda t: 0.1 * t)
    heat = lambda u, x, t: [diff(u, t) - 0.3 * diff(u, x, order=2)]          # noqa: E731
    rs = np.random.RandomState(8)
    xs, ts = rs.rand(40), rs.rand(40)
    fp = CpuFusedProblem([net], [C.IBVP1D(**kw)], heat, 2)
    assert len(fp.tp.nets) == 2 and fp.tp.nets[0].module is fp.tp.nets[1].module and fp.tp.nets[0].skip is not None
    fp.gradbuf.zero_()
    sumsq, _ = fp.residual_grad([torch.tensor(xs), torch.tensor(ts)])
    loss = float(sumsq) / 40
    got = {name: p.grad.clone() for name, p in net.named_parameters()}
    for p in net.parameters():
        p.grad = None
    cols = [torch.tensor(v).reshape(-1, 1).requires_grad_(True) for v in (xs, ts)]
    u = oracle.IBVP1D(**kw).enforce(net, *cols)
    ref = ((oracle.diff(u, cols[1]) - 0.3 * oracle.diff(u, cols[0], order=2)) ** 2).mean()
    ref.backward()
    assert abs(loss - float(ref.detach())) <= 1e-6 * float(ref.detach())
    for name, p in net.named_parameters():
        np.testing.assert_allclose(got[name].numpy(), p.grad.numpy(), rtol=2e-6, atol=1e-9, err_msg=name)
