"""Randomised check of the tracer: user-style expressions built from the SAME Python source are evaluated once on torch
tensors with autograd (what the reference does, neurodiffeq.py:6-34) and once on traced symbols (symbolic forward
differentiation -> lowered bytecode -> numpy interpreter).  Values, first/second/mixed derivatives w.r.t. the coordinates
and the reverse-mode adjoints w.r.t. coordinate-independent leaves must agree to rounding."""
import numpy as np
import pytest
import torch

from neurodiffeq_b200 import symbolic as S
from neurodiffeq_b200 import diff

UNARY = [torch.sin, torch.cos, torch.tanh, torch.exp, lambda a: a * a, lambda a: 1.0 / (1.5 + a * a),
         lambda a: torch.sqrt(1.0 + a * a), lambda a: torch.log(2.0 + a * a), torch.atan, lambda a: abs(a) * a]


def random_expression(rs, leaves, depth):
    """A random expression tree over `leaves` (callables returning the leaf in the current domain)."""
    if depth == 0 or rs.rand() < 0.15:
        k = rs.randint(len(leaves) + 1)
        if k == len(leaves):
            c = float(np.round(rs.uniform(-2, 2), 3))
            return lambda env: c
        return lambda env: env[k]
    kind = rs.randint(4)
    a = random_expression(rs, leaves, depth - 1)
    if kind == 0:
        f = UNARY[rs.randint(len(UNARY))]
        return lambda env: f(a(env) + 0.0 * env[0])       # keep the argument a tensor / symbol
    b = random_expression(rs, leaves, depth - 1)
    if kind == 1:
        return lambda env: a(env) + b(env)
    if kind == 2:
        return lambda env: a(env) * b(env)
    return lambda env: a(env) - 0.5 * b(env)


@pytest.mark.parametrize("seed", range(12))
def test_traced_expression_and_derivatives_match_autograd(seed):
    rs = np.random.RandomState(seed)
    n = 17
    xs = rs.uniform(-1, 1, size=(2, n))
    ys = rs.uniform(-1, 1, size=(2, n))                   # two "network output" leaves (value channels)
    expr = random_expression(rs, [None] * 4, depth=4)

    # --- torch: coordinates and jet leaves are tensors; d/dx via autograd ---
    tx = [torch.tensor(x, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for x in xs]
    ty = [torch.tensor(y, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for y in ys]
    val_t = expr(tx + ty) + 0.0 * tx[0]
    d0_t = torch.autograd.grad(val_t.sum(), tx[0], create_graph=True, allow_unused=True)[0]
    d0_t = torch.zeros_like(tx[0]) if d0_t is None else d0_t
    d01_t = torch.autograd.grad(d0_t.sum(), tx[1], create_graph=True, allow_unused=True)[0] if d0_t.requires_grad else None
    d01_t = torch.zeros_like(tx[0]) if d01_t is None else d01_t
    d00_t = torch.autograd.grad(d0_t.sum(), tx[0], create_graph=True, allow_unused=True)[0] if d0_t.requires_grad else None
    d00_t = torch.zeros_like(tx[0]) if d00_t is None else d00_t
    r_t = val_t + 0.3 * d0_t - 0.7 * d01_t + 0.2 * d00_t   # a "residual" mixing the derivatives
    adj_t = torch.autograd.grad(r_t.sum(), ty, allow_unused=True)
    adj_t = [torch.zeros_like(ty[0]) if a is None else a for a in adj_t]

    # --- symbols: same Python, diff() is exact symbolic differentiation ---
    g = S.Graph()
    g.n_sampled = 2
    sx = [g.coord(0), g.coord(1)]
    sy = [g.rbar(0), g.rbar(1)]                            # leaves that do not depend on the coordinates (like ty above)
    val_s = g.lift(expr(sx + sy) + 0.0 * sx[0])
    d0_s = diff(val_s, sx[0])
    r_s = val_s + 0.3 * d0_s - 0.7 * diff(d0_s, sx[1]) + 0.2 * diff(val_s, sx[0], order=2)
    adj = S.reverse_gradients([(g.lift(r_s), g.const(1.0))], wrt_filter=lambda node: node.op == "rbar")
    rows = {sy[0]: 0, sy[1]: 1}
    outs = [(S.OP_ST_U, 0, val_s), (S.OP_ST_R, 0, g.lift(r_s))]
    outs += [(S.OP_ST_SEED, rows[leaf], e) for leaf, e in adj.items()]
    prog = S.lower(outs, lambda net, o, c: o)
    u, r, seed_rows = S.evaluate_program(prog, xs, np.zeros((1, n)), rbar=ys, n_u=1, n_r=1, n_seed=2)

    scale = 1.0 + np.abs(r_t.detach().numpy()[:, 0])
    # constants of the lowered program are float32 immediates unless exactly representable -> 1e-6 level agreement
    np.testing.assert_allclose(u[0], val_t.detach().numpy()[:, 0], rtol=2e-6, atol=2e-6)
    assert np.max(np.abs(r[0] - r_t.detach().numpy()[:, 0]) / scale) < 5e-6
    for k in range(2):
        ref = adj_t[k].detach().numpy()[:, 0]
        assert np.max(np.abs(seed_rows[k] - ref) / (1.0 + np.abs(ref))) < 5e-6


@pytest.mark.parametrize("seed", range(24))
def test_random_pde_residuals_value_loss_and_gradient_match_autograd(seed):
    """End to end through the tracer, the channel scheme (incl. mixed partials by polarisation and the combined
    second-order channel when the residual happens to be affine in u_xx, u_yy), the lowered programs and the numpy mirror
    of the kernels -- against plain autograd on the same small network, for random residual expressions over
    {u, u_x, u_y, u_xx, u_yy, u_xy, x, y}."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_engine import CpuFusedProblem
    from neurodiffeq_b200.conditions import NoCondition
    from neurodiffeq_b200.networks import FCNN
    rs = np.random.RandomState(100 + seed)
    torch.manual_seed(seed)
    net = FCNN(2, 1, hidden_units=(7, 6)).double()
    use_mixed = rs.rand() < 0.5
    shape = random_expression(rs, [None] * 8, depth=3)
    affine = rs.rand() < 0.4          # sometimes an affine-in-second-derivatives residual: exercises the combined channel

    def residual(u, x, y):
        ux, uy = diff(u, x), diff(u, y)
        uxx, uyy = diff(u, x, order=2), diff(u, y, order=2)
        if affine:
            return [(1.0 + x * x) * uxx + torch.exp(-y) * uyy + shape([u, ux, uy, x, y, x, y, u])]
        uxy = diff(ux, y) if use_mixed else ux * uy
        return [shape([u, ux, uy, uxx, uyy, uxy, x, y])]

    n = 23
    xs, ys = rs.uniform(-1, 1, n), rs.uniform(-1, 1, n)
    fp = CpuFusedProblem([net], [NoCondition()], residual, 2)
    if affine:
        assert fp.tp.wl == 2                      # the tracer proved affinity: one weighted second-order channel
    fp.gradbuf.zero_()
    sumsq, r = fp.residual_grad([torch.tensor(xs), torch.tensor(ys)], want_residual=True)
    got_grads = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    cx = [torch.tensor(v).reshape(-1, 1).requires_grad_(True) for v in (xs, ys)]
    u = net(torch.cat(cx, dim=1))
    res = residual(u, *cx)[0] + 0.0 * u
    loss = (res ** 2).mean()
    loss.backward()
    scale = 1.0 + np.abs(res.detach().numpy()[:, 0])
    assert np.max(np.abs(r.numpy()[0] - res.detach().numpy()[:, 0]) / scale) < 1e-5
    assert abs(float(sumsq) / n - float(loss.detach())) <= 1e-5 * (1.0 + float(loss.detach()))
    gn = np.sqrt(sum(float((p.grad ** 2).sum()) for p in net.parameters()))
    dn = np.sqrt(sum(float(((a - p.grad) ** 2).sum()) for a, p in zip(got_grads, net.parameters())))
    assert dn <= 1e-5 * (1.0 + gn)
