"""Host logic: tracing the reference-style problem definitions -> programs; algebra of the kernels (numpy mirror)
against golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

import workloads
from conftest import load_golden
from oracle import jet_numpy


from helpers import product_namespace  # noqa: E402


def params_per_instance(tp, flat_params):
    """golden params are in state_dict order over the distinct modules; instances of one module share them"""
    by_module, it, per_net = {}, iter(flat_params), []
    for nd in tp.nets:
        if id(nd.module) not in by_module:
            by_module[id(nd.module)] = [next(it) for _ in range(len(nd.parameters()))]
        per_net.append(by_module[id(nd.module)])
    return per_net


def trace(key):
    from neurodiffeq_b200.tracing import TracedProblem
    wl = workloads.build(product_namespace(), key)
    nets, conds = wl.make_nets(), wl.make_conditions()
    tp = TracedProblem(nets, conds, workloads.bundle_eq_wrapper(wl), len(wl.coord_names))
    return wl, nets, conds, tp


EXPECTED_CHANNELS = {"c1": (1, 0), "c2": (2, 2), "c3": (2, 1), "c4": (3, 3), "c5": (1, 0),
                     # Neumann ends: the network is also evaluated at a constant coordinate; heat: x, t, boundary, t+boundary
                     "x1": (4, 4), "x2": (4, 4), "x3": (2, 1), "x4": (2, 1), "x5": (2, 1), "x6": (1, 1), "x8": (4, 4), "x9": (2, 1),
                     "x7": (1, 0)}   # EnsembleCondition: one 2-output network, the function is an (N, 2) block


@pytest.mark.parametrize("key", workloads.NAMES + workloads.EXTRA_NAMES)
def test_traced_problem_matches_reference_golden(key):
    wl, nets, conds, tp = trace(key)
    assert (tp.scheme.n1, tp.scheme.n2) == EXPECTED_CHANNELS[key]
    gold = load_golden(wl.name)
    out = jet_numpy.run_traced(tp, params_per_instance(tp, gold["params"]), gold["coords"])
    rms = np.sqrt((gold["residual"] ** 2).mean())
    np.testing.assert_allclose(out["u"], gold["u"], rtol=1e-10, atol=1e-12)
    assert np.abs(out["residual"] - gold["residual"]).max() <= 1e-9 * rms
    assert abs(out["loss"] - gold["loss"]) <= 1e-10 * gold["loss"]
    gn = np.sqrt(sum((g ** 2).sum() for g in gold["grads"]))
    dn = np.sqrt(sum(((g - h.reshape(g.shape)) ** 2).sum() for g, h in zip(gold["grads"], out["grads"])))
    assert dn <= 1e-9 * gn
    print(key, "eval prog", len(tp.prog_eval), "slots", tp.prog_eval.n_slots, "| train prog", len(tp.prog_train),
          "slots", tp.prog_train.n_slots)


@pytest.mark.parametrize("key,wl,channels", [("c2", 2, 4), ("c4", 3, 5), ("c3", 0, 4), ("c5", 0, 2), ("x1", 4, 6),
                                             ("x2", 4, 6), ("x8", 4, 6), ("x5", 0, 4)])
def test_combined_second_order_channel(key, wl, channels):
    """Residuals affine in the pure second derivatives with coordinate-only coefficients are carried as ONE weighted
    channel (forward-Laplacian style): C2 5 -> 4 channels, C4 7 -> 5; values and gradients are unchanged."""
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200.engine import pad_scheme
    wl_, nets, conds, _ = trace(key)
    tp = TracedProblem(nets, conds, workloads.bundle_eq_wrapper(wl_), len(wl_.coord_names), pad_scheme=pad_scheme,
                       combine_seconds=lambda a, b: (a, b) in ((2, 2), (3, 3), (4, 4)))
    assert tp.wl == wl and tp.n_channels == channels
    gold = load_golden(wl_.name)
    out = jet_numpy.run_traced(tp, params_per_instance(tp, gold["params"]), gold["coords"])
    rms = np.sqrt((gold["residual"] ** 2).mean())
    assert np.abs(out["residual"] - gold["residual"]).max() <= 1e-9 * rms
    gn = np.sqrt(sum((g ** 2).sum() for g in gold["grads"]))
    dn = np.sqrt(sum(((g - h.reshape(g.shape)) ** 2).sum() for g, h in zip(gold["grads"], out["grads"])))
    assert dn <= 1e-9 * gn


def test_combined_channel_is_refused_when_not_affine():
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200.networks import FCNN
    from neurodiffeq_b200.conditions import NoCondition
    from neurodiffeq_b200.tracing import TracedProblem
    net = FCNN(2, 1, hidden_units=(8,))
    always = lambda a, b: True  # noqa: E731
    # u * u_xx: the coefficient of the second derivative depends on the network -> separate channels stay
    tp = TracedProblem([net], [NoCondition()], lambda u, x, y: [u * diff(u, x, order=2) + diff(u, y, order=2)], 2,
                       combine_seconds=always)
    assert tp.wl == 0
    # sin(u_xx): not affine
    tp = TracedProblem([net], [NoCondition()], lambda u, x, y: [torch.sin(diff(u, x, order=2)) + diff(u, y, order=2)], 2,
                       combine_seconds=always)
    assert tp.wl == 0
    # two equations that need different combinations of the same net's second derivatives
    tp = TracedProblem([net], [NoCondition()],
                       lambda u, x, y: [diff(u, x, order=2) + diff(u, y, order=2), diff(u, x, order=2) - diff(u, y, order=2)],
                       2, combine_seconds=always)
    assert tp.wl == 0


def test_mixed_partials_by_polarisation():
    """u_xy is carried as (D_{x+y}^2 - D_x^2 - D_y^2)/2; checked against autograd on a random FCNN."""
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200.networks import FCNN
    from neurodiffeq_b200.conditions import NoCondition
    from neurodiffeq_b200.tracing import TracedProblem
    torch.manual_seed(3)
    net = FCNN(2, 1, hidden_units=(16, 16)).double()
    eq = lambda u, x, y: [diff(diff(u, x), y) + diff(u, x, order=2) * y - diff(u, y)]  # noqa: E731
    tp = TracedProblem([net], [NoCondition()], eq, 2)
    assert tp.scheme.n2 == 3 and tp.scheme.n1 == 3
    coords = np.random.RandomState(0).rand(2, 64)
    params = [[p.detach().numpy() for p in net.parameters()]]
    out = jet_numpy.run_traced(tp, params, coords)
    x, y = (torch.tensor(c).reshape(-1, 1).requires_grad_(True) for c in coords)
    u = net(torch.cat([x, y], 1))
    r = eq(u, x, y)[0]
    loss = (r ** 2).mean()
    loss.backward()
    np.testing.assert_allclose(out["residual"][0], r.detach().numpy()[:, 0], rtol=1e-9, atol=1e-11)
    for g, p in zip(out["grads"], net.parameters()):
        ref = np.zeros(p.shape) if p.grad is None else p.grad.numpy()  # output bias: no gradient at all
        np.testing.assert_allclose(g.reshape(p.shape), ref, rtol=1e-8, atol=1e-11)


def test_unused_coordinate_gives_zero_and_order3_raises():
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200.networks import FCNN
    from neurodiffeq_b200.conditions import NoCondition
    from neurodiffeq_b200.tracing import TracedProblem
    net = FCNN(1, 1, hidden_units=(8,))
    tp = TracedProblem([net], [NoCondition()], lambda u, t: [diff(u, t) + diff(t * t, t) + diff(3.0 * t, t, order=2)], 1)
    assert (tp.scheme.n1, tp.scheme.n2) == (1, 0)
    with pytest.raises(NotImplementedError):
        TracedProblem([net], [NoCondition()], lambda u, t: [diff(u, t, order=3)], 1)
    with pytest.raises(TypeError):
        TracedProblem([net], [NoCondition()], lambda u, t: [u if u > 0 else -u], 1)


@pytest.mark.parametrize("key", ["c1", "c5"])
def test_h1_loss_is_the_mean_square_of_augmented_rows(key):
    """'h1' (reference losses.py:17-20) = mean square of [residual | grad(residual, *coords)], where grad differentiates
    the SUM of the residual columns: traced as extra residual rows, it is the plain fused L2 path.  Checked against
    autograd (oracle) for loss and parameter gradients on a first-order system (second-order jets suffice)."""
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200.losses import h1_rows, _losses
    from oracle import reference_port as oracle
    wl = workloads.build(product_namespace(), key)
    torch.manual_seed(4)
    nets, conds = wl.make_nets(), wl.make_conditions()
    n_funcs = len(conds)
    tp = TracedProblem(nets, conds, h1_rows(workloads.bundle_eq_wrapper(wl), n_funcs), len(wl.coord_names))
    assert tp.n_eq == wl.n_eq + len(wl.coord_names)
    from helpers import get_params
    params = get_params(nets)
    coords = workloads.sample_coords(wl, 200, seed=8)
    out = jet_numpy.run_traced(tp, params_per_instance(tp, params), coords)
    # autograd: the reference's loss function on the oracle's residual matrix
    owl = workloads.build(oracle.NAMESPACE, key)
    onets, oconds = owl.make_nets(), owl.make_conditions()
    oracle.load_params(onets, params, dtype=torch.float64)
    cols = [torch.as_tensor(c, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for c in coords]
    funcs = [c.enforce(n, *cols) for n, c in zip(onets, oconds)]
    res = torch.cat(workloads.bundle_eq_wrapper(owl)(*funcs, *cols), dim=1)
    g = oracle.grad(res, *cols)
    loss = (torch.cat([res, *g], dim=1) ** 2).mean()
    loss_product_eager = float(_losses["h1"](res, funcs, cols).detach())   # the product's own loss function on tensors
    loss.backward()
    loss = float(loss.detach())
    assert abs(loss_product_eager - loss) <= 1e-12 * loss
    assert abs(out["loss"] - loss) <= 1e-10 * loss
    ref_grads = [p.grad.numpy() for m in oracle.distinct_modules(onets) for p in m.parameters()]
    gn = np.sqrt(sum((a ** 2).sum() for a in ref_grads))
    dn = np.sqrt(sum(((a - b.reshape(a.shape)) ** 2).sum() for a, b in zip(ref_grads, out["grads"])))
    assert dn <= 1e-9 * gn
