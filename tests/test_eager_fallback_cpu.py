"""what the tracer / kernels refuse falls back to the autograd path (neurodiffeq_b200/eager.py) with ONE
warning, and trains exactly like the reference closure.  CPU: the fused engine is the float64 stand-in (its tracer is the
product's, so the refusals are the real ones); the fallback itself is the product's EagerProblem on CPU tensors."""
import warnings

import numpy as np
import pytest
import torch

import workloads
from cpu_engine import CpuFusedProblem
from helpers import get_params
from test_solvers_gpu import make_solver, oracle_training
from test_losses_gpu import oracle_training_with_loss


@pytest.fixture(autouse=True)
def cpu_engine(monkeypatch):
    import neurodiffeq_b200.solvers as S
    import neurodiffeq_b200.eager as E
    monkeypatch.setattr(S, "FusedProblem", CpuFusedProblem)
    monkeypatch.setattr(E, "_WARNED", set())
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


@pytest.mark.parametrize("key", workloads.FALLBACK_NAMES)
def test_refused_problem_trains_on_the_autograd_path(key):
    n, epochs = 64, 4
    with pytest.warns(RuntimeWarning, match="falling back to the autograd path"):
        wl, solver, nets, coords_np = make_solver(key, n, device="cpu")
    assert getattr(solver.problem, "is_eager", False) and solver.problem.reason
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training(key, params0, coords_np, epochs)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    np.testing.assert_allclose(solver.metrics_history["valid_loss"][:-1], ref_losses[1:], rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    # solutions and residuals come from the same path
    pts = [torch.linspace(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), 9) * (1.0 + 0.01 * i)
           for i, (lo, hi) in enumerate(wl.coord_ranges)]
    u = solver.get_solution(best=False)(*pts)
    r = solver.get_residuals(*pts, best=False)
    assert u.shape == (9,) and r.shape == (9,)
    from oracle import reference_port as oracle
    owl = workloads.build(oracle.NAMESPACE, key)
    onets, oconds = owl.make_nets(), owl.make_conditions()
    oracle.load_params(onets, get_params(nets), dtype=torch.float64)
    cols = [p.reshape(-1, 1).requires_grad_(True) for p in pts]
    uo = oconds[0].enforce(onets[0], *cols)
    ro = owl.diff_eqs(uo, *cols)[0]
    np.testing.assert_allclose(u.detach().numpy(), uo.detach().numpy().ravel(), rtol=1e-10)
    np.testing.assert_allclose(r.detach().numpy(), ro.detach().numpy().ravel(), rtol=1e-8, atol=1e-10)


def test_h1_on_a_second_order_problem_falls_back():
    """reference losses.py:17-20: the h1 loss differentiates the residual once more -> order 3 on a second-order BVP."""
    key, n, epochs = "x6", 48, 3
    with pytest.warns(RuntimeWarning, match="falling back to the autograd path"):
        wl, solver, nets, coords_np = make_solver(key, n, loss_fn="h1", device="cpu")
    assert solver.problem.is_eager
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_with_loss(key, params0, coords_np, epochs, "h1")
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    r = solver.get_residuals(torch.linspace(0.1, 0.9, 5), best=False)      # the user's equation only, not the h1 rows
    assert r.shape == (5,)


def test_custom_loss_and_device_loop_on_the_autograd_path():
    def loss_fn(residual, funcs, coords):
        return (residual ** 2).mean() + 0.1 * (funcs[0] ** 2).mean()

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        wl, solver, nets, coords_np = make_solver("y1", 40, loss_fn=loss_fn, device="cpu")
    params0 = get_params(nets)
    solver.fit(3, tqdm_file=None)
    from helpers import oracle_training_custom
    ref_losses, ref_params = oracle_training_custom("y1", params0, coords_np, 3, loss_fn)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    solver.device_loop = True
    assert "autograd path" in solver._device_loop_blocker()


def test_one_warning_per_reason_and_real_errors_propagate():
    with pytest.warns(RuntimeWarning) as rec:
        make_solver("y2", 16, device="cpu")
        make_solver("y2", 16, device="cpu")
    assert sum("falling back" in str(w.message) for w in rec) == 1

    from neurodiffeq_b200 import solvers as S
    from neurodiffeq_b200.generators import PredefinedGenerator
    nd = workloads.product_namespace()
    wl = workloads.build(nd, "y2")
    gen = PredefinedGenerator(np.linspace(0, 1, 8).astype(np.float32))

    def broken(u, t):
        return [nd.diff(u, t) + undefined_name]   # noqa: F821

    with pytest.raises(NameError):
        S.Solver1D(broken, wl.make_conditions(), nets=wl.make_nets(), train_generator=gen, valid_generator=gen, device="cpu")


@pytest.mark.parametrize("kind", ["swish", "aptx", "monomial"])
def test_reference_modules_without_jet_rules_train_like_plain_autograd(kind):
    """Swish / APTx / MonomialNN (reference networks.py:109-208) are refused by the tracer; on the autograd path the solver
    must do exactly what a hand-written torch loop over the same modules does."""
    from copy import deepcopy
    from neurodiffeq_b200 import solvers as S, diff
    from neurodiffeq_b200.conditions import IVP
    from neurodiffeq_b200.generators import PredefinedGenerator
    from neurodiffeq_b200.networks import FCNN, Swish, APTx, MonomialNN
    torch.manual_seed(3)
    if kind == "monomial":
        net = torch.nn.Sequential(MonomialNN(degrees=(1, 2, 3)), torch.nn.Linear(3, 1))   # monomial features -> u
    else:
        net = FCNN(n_input_units=1, n_output_units=1, hidden_units=(8, 8), actv=Swish if kind == "swish" else APTx)
    twin = deepcopy(net)
    t_np = np.linspace(0.05, 1.5, 24)
    gen = PredefinedGenerator(t_np)
    eq = lambda u, t: [diff(u, t) + 2.0 * u]   # noqa: E731
    with pytest.warns(RuntimeWarning, match="falling back to the autograd path"):
        solver = S.Solver1D(eq, [IVP(t_0=0.0, u_0=1.0)], nets=[net], train_generator=gen, valid_generator=gen,
                            n_batches_valid=0, device="cpu")
    solver.fit(3, tqdm_file=None)
    opt, cond, losses = torch.optim.Adam(twin.parameters(), lr=1e-3), IVP(t_0=0.0, u_0=1.0), []
    for _ in range(3):
        opt.zero_grad()
        t = torch.as_tensor(t_np).reshape(-1, 1).requires_grad_(True)
        u = cond.enforce(twin, t)
        loss = (eq(u, t)[0] ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    np.testing.assert_allclose(solver.metrics_history["train_loss"], losses, rtol=5e-7)
    for a, b in zip(net.parameters(), twin.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=1e-9, atol=1e-12)


def test_live_solution_follows_training_on_the_autograd_path():
    """get_solution(copy=False, best=False) hands out the LIVE networks (reference solvers.py:606-640): evaluated again
    after more training it must show the new parameters."""
    import warnings as w
    with w.catch_warnings():
        w.simplefilter("ignore", RuntimeWarning)
        wl, solver, nets, coords_np = make_solver("y2", 32, device="cpu")
    solver.fit(2, tqdm_file=None)
    live = solver.get_solution(copy=False, best=False)
    t = torch.linspace(0.1, 1.9, 7)
    u1 = live(t).clone()
    solver.fit(3, tqdm_file=None)
    u2 = live(t)
    fresh = solver.get_solution(copy=True, best=False)(t)
    np.testing.assert_allclose(u2.detach().numpy(), fresh.detach().numpy(), rtol=1e-12)
    assert float((u2 - u1).abs().max()) > 0
    solver.fit(1, tqdm_file=None)               # the solver re-adopts its parameters after the solution borrowed them
    assert np.isfinite(solver.metrics_history["train_loss"][-1])


@pytest.mark.parametrize("key", workloads.FALLBACK_NAMES)
def test_autograd_path_matches_reference_golden(key):
    """The fallback engine itself (forward / residual_grad through the FusedProblem interface) against golden vectors made by
    the UNMODIFIED reference (tests/golden/generate.py) -- the same check the fused kernels get for C1..C5 / x1..x9."""
    from conftest import load_golden
    from neurodiffeq_b200.eager import EagerProblem
    wl = workloads.build(workloads.product_namespace(), key)
    gold = load_golden(wl.name)
    nets, conds = wl.make_nets(), wl.make_conditions()
    workloads.set_params(nets, gold["params"])
    ep = EagerProblem(nets, conds, wl.diff_eqs, len(wl.coord_names), device="cpu", reason="test")
    coords = [torch.from_numpy(c.astype(np.float64)) for c in gold["coords"]]
    u, r, s = ep.forward(coords, want_sumsq=True)
    n = gold["coords"].shape[1]
    np.testing.assert_allclose(u.numpy(), gold["u"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(r.numpy(), gold["residual"], rtol=1e-10, atol=1e-11)
    assert abs(float(s) / (n * ep.n_eq) - gold["loss"]) <= 1e-12 * gold["loss"]
    ep.gradbuf.zero_()
    s2, r2 = ep.residual_grad(coords, want_residual=True)
    assert abs(float(s2) / (n * ep.n_eq) - gold["loss"]) <= 1e-12 * gold["loss"]
    for g, h in zip(ep.grads_as_list(), gold["grads"]):
        np.testing.assert_allclose(g, h, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(h).max()))
    ep.residual_grad(coords)                      # accumulates like loss.backward()
    for g, h in zip(ep.grads_as_list(), gold["grads"]):
        np.testing.assert_allclose(g, 2.0 * h, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(h).max()))


def test_planner_refusal_becomes_a_fallback_reason():
    """engine.planner_refusal: return code -2 of pj_sizes (e.g. hidden width > PJ_MAX_WIDTH) -> text, anything else -> None
    (other failures keep raising where they always did).  Checked with a stand-in library object: no GPU here."""
    from neurodiffeq_b200 import engine as E

    class Lib:
        def __init__(self, rc):
            self.rc = rc

        def pj_sizes(self, spec, n, out):
            return self.rc

        def pj_last_error(self):
            return b"net 0: hidden width 256 not in 1..128"

    spec = E.PjSpec()
    assert E.planner_refusal(Lib(0), spec, "cpu") is None
    assert E.planner_refusal(Lib(-4), spec, "cpu") is None
    assert "hidden width 256" in E.planner_refusal(Lib(-2), spec, "cpu")
