"""Host logic of the fused solvers on the CPU: the engine is replaced by a float64 stand-in built on the numpy mirror of
the kernels (tests/cpu_engine.py), everything else -- tracing, epoch loop, named / custom losses, best-network tracking,
solutions and residuals -- is the product's code.  Training must follow the oracle (autograd + torch Adam, float64) to
rounding level.  The GPU suite repeats these through the real kernels at fp32 tolerances."""
import numpy as np
import pytest
import torch

import workloads
from cpu_engine import CpuFusedProblem
from helpers import get_params, oracle_training_custom, oracle_training_lbfgs
from test_solvers_gpu import make_solver, oracle_training
from test_losses_gpu import oracle_training_with_loss


@pytest.fixture(autouse=True)
def cpu_engine(monkeypatch):
    import neurodiffeq_b200.solvers as S
    monkeypatch.setattr(S, "FusedProblem", CpuFusedProblem)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


@pytest.mark.parametrize("key", ["c1", "c2", "c4", "c5", "x1", "x4", "x5", "x7", "x8", "x9"])
def test_fit_tracks_oracle_adam_cpu(key):
    n, epochs = 160, 4
    wl, solver, nets, coords_np = make_solver(key, n)
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training(key, params0, coords_np, epochs)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    np.testing.assert_allclose(solver.metrics_history["valid_loss"][:-1], ref_losses[1:], rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-11)
    assert solver.global_epoch == epochs and solver.lowest_loss == min(solver.metrics_history["valid_loss"])


@pytest.mark.parametrize("loss_name", ["l1", "infinity", "h1", "h1 semi"])
def test_named_losses_cpu(loss_name):
    key, n, epochs = "c1", 120, 4
    wl, solver, nets, coords_np = make_solver(key, n, loss_fn=loss_name)
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_with_loss(key, params0, coords_np, epochs, loss_name)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-11)
    r = solver.get_residuals(torch.linspace(0.5, 2.0, 7), best=False)
    assert isinstance(r, list) and len(r) == wl.n_eq and r[0].shape == (7,)


def test_loss_name_errors_cpu():
    with pytest.raises(KeyError):
        make_solver("c1", 16, loss_fn="l3")


def test_custom_loss_callable_and_solution_cpu():
    def weighted(residual, funcs, coords):          # (residual, funcs, coords) -> scalar, reference solvers.py:66-79
        return (residual ** 2 * (1.0 + coords[0] ** 2)).mean()

    wl, solver, nets, coords_np = make_solver("x3", 100, loss_fn=weighted)
    l0 = None
    solver.fit(3, tqdm_file=None)
    hist = solver.metrics_history["train_loss"]
    assert len(hist) == 3 and all(np.isfinite(hist))
    sol = solver.get_solution(best=False)
    x = torch.tensor([0.0, 0.5, 1.0])
    u = sol(x, to_numpy=True)
    assert u.shape == (3,) and abs(u[0] - 1.0) < 1e-12     # Dirichlet end of x3: u(0) = 1 for any weights
    del l0


def test_solution_residuals_metrics_and_early_stop_cpu():
    """The solver API surface of the reference (solvers.py:443-497, 606-720) on the stand-in engine: solution / residual
    evaluation with reshaping, `best` networks, metrics histories, callbacks with early stopping, multi-batch epochs."""
    wl, solver, nets, coords_np = make_solver("c2", 100, metrics={"mean_u": lambda u, x, y: u.mean()})
    calls = []

    def stopper(s):
        calls.append(s.local_epoch)
        if s.local_epoch == 2:
            s._stop_training = True

    solver.fit(10, callbacks=[stopper], tqdm_file=None)
    assert calls == [1, 2] and solver.global_epoch == 2
    assert len(solver.metrics_history["train__mean_u"]) == 2 and len(solver.metrics_history["valid__mean_u"]) == 2
    xs, ys = np.linspace(0, 1, 5), np.linspace(0, 1, 4)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    u = solver.get_solution(best=False)(X, Y, to_numpy=True)
    assert u.shape == X.shape
    np.testing.assert_allclose(u[0, :], np.sin(np.pi * ys), atol=1e-6)       # Dirichlet data met by construction
    np.testing.assert_allclose(u[-1, :], 0, atol=1e-6)
    r = solver.get_residuals(X, Y, to_numpy=True, best=False)
    assert r.shape == X.shape and np.isfinite(r).all()
    rb = solver.get_residuals(torch.tensor(X), torch.tensor(Y), best=True)
    assert isinstance(rb, torch.Tensor) and rb.shape == X.shape
    assert solver.get_solution(best=True)(X, Y, to_numpy=True, no_reshape=True).shape == (X.size, 1)
    # two batches per epoch: gradients add up (reference solvers.py:360-362)
    wl, s2, nets2, _ = make_solver("c2", 60, n_batches_train=2)
    s2.run_train_epoch()
    g2 = s2.problem.grad.clone()
    wl, s1, nets1, _ = make_solver("c2", 60)
    s1.run_train_epoch()
    assert torch.allclose(g2, 2 * s1.problem.grad, rtol=1e-9, atol=1e-12)


def test_ensemble_solution_is_a_block_cpu():
    wl, solver, nets, coords_np = make_solver("x7", 80)
    solver.fit(2, tqdm_file=None)
    t = torch.linspace(0.0, 1.0, 6)
    uv = solver.get_solution(best=False)(t, to_numpy=True)
    assert uv.shape == (6, 2)
    np.testing.assert_allclose(uv[0], [0.0, 1.0], atol=1e-12)              # the two initial values
    r = solver.get_residuals(t, best=False)
    assert isinstance(r, list) and len(r) == 2 and r[0].shape == (6,)


def test_lbfgs_closure_mode_cpu():
    """LBFGS (reference solvers.py:398-400): one step(closure) per batch; the closure re-evaluates loss and gradient.
    Same optimizer, same closure semantics on the oracle (autograd) -> same parameters."""
    key, n, epochs = "x6", 60, 3
    wl = workloads.build(__import__("helpers").product_namespace(), key)
    torch.manual_seed(0)
    nets = wl.make_nets()
    opt = torch.optim.LBFGS([p for m in nets for p in m.parameters()], lr=0.5, max_iter=4, history_size=5)
    import neurodiffeq_b200.solvers as Sv
    from neurodiffeq_b200.generators import PredefinedGenerator
    coords_np = workloads.sample_coords(wl, n, seed=21)
    gen = PredefinedGenerator(*[c for c in coords_np])
    params0 = get_params(nets)
    solver = Sv.Solver1D(wl.diff_eqs, wl.make_conditions(), nets=nets, train_generator=gen, valid_generator=gen,
                         n_batches_valid=1, optimizer=opt)
    solver.fit(epochs, tqdm_file=None)

    ref_losses, ref_params = oracle_training_lbfgs(key, params0, coords_np, epochs, lr=0.5, max_iter=4, history_size=5)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=1e-5)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-8)
    assert len(solver.metrics_history["valid_loss"]) == epochs


def test_loss_that_depends_on_the_functions_cpu():
    """loss_fn(residual, funcs, coords) may look at the functions (reference solvers.py:66-79): dL/du reaches the
    parameters through the same traced program as dL/dr."""
    def loss_fn(residual, funcs, coords):
        u, v = funcs
        return (residual ** 2).mean() + 0.3 * ((u - 1.0) ** 2).mean() + 0.1 * (u * v * coords[0]).mean()

    key, n, epochs = "c1", 90, 4
    wl, solver, nets, coords_np = make_solver(key, n, loss_fn=loss_fn)
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_custom(key, params0, coords_np, epochs, loss_fn)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-11)


def test_overridden_additional_loss_cpu():
    """`additional_loss` (reference solvers.py:587-604) as a subclass hook, here a penalty on the solution's mean."""
    import neurodiffeq_b200.solvers as Sv
    from neurodiffeq_b200.generators import PredefinedGenerator

    class Penalised(Sv.Solver1D):
        def additional_loss(self, residual, funcs, coords):
            return 0.5 * funcs[0].mean() ** 2

    key, n, epochs = "x6", 70, 4
    wl = workloads.build(__import__("helpers").product_namespace(), key)
    torch.manual_seed(0)
    nets = wl.make_nets()
    coords_np = workloads.sample_coords(wl, n, seed=21)
    gen = PredefinedGenerator(*[c for c in coords_np])
    params0 = get_params(nets)
    solver = Penalised(wl.diff_eqs, wl.make_conditions(), nets=nets, train_generator=gen, valid_generator=gen, n_batches_valid=1)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_custom(
        key, params0, coords_np, epochs, lambda r, f, x: (r ** 2).mean() + 0.5 * f[0].mean() ** 2)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-7)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("capturable", [False, True])
def test_flat_adam_equals_torch_adam_cpu(capturable):
    """optim.FlatAdam on the engine's flat buffers takes the same steps as torch.optim.Adam on the parameter views; the
    capturable variant keeps step count / learning rate / bias corrections in device tensors (graph-recordable)."""
    from neurodiffeq_b200.optim import FlatAdam
    key, n, epochs = "c2", 50, 6
    wl, s_ref, nets_ref, coords_np = make_solver(key, n)
    p0 = get_params(nets_ref)
    s_ref.fit(epochs, tqdm_file=None)
    wl, s_flat, nets_flat, _ = make_solver(key, n)
    from helpers import set_params
    set_params(nets_flat, p0)
    s_flat.optimizer = FlatAdam.for_solver(s_flat, lr=1e-3, capturable=capturable)
    s_flat.fit(epochs, tqdm_file=None)
    np.testing.assert_allclose(s_flat.metrics_history["train_loss"], s_ref.metrics_history["train_loss"], rtol=1e-10)
    for a, b in zip(get_params(nets_flat), get_params(nets_ref)):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)
    for s_ in (s_flat, s_ref):                          # schedulers / callbacks edit param_groups as with any optimizer
        s_.optimizer.param_groups[0]["lr"] = 5e-4
        s_.fit(2, tqdm_file=None)
    for a, b in zip(get_params(nets_flat), get_params(nets_ref)):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)


def test_training_converges_to_the_analytic_solution_cpu():
    """End to end, like the reference's tests/test_ode.py: u' = -u, u(0) = 1 on [0, 2] trained with Adam reaches the
    analytic solution exp(-t) (loose bound; the point is that value, residual, gradient, optimizer and best-network
    bookkeeping work together)."""
    import neurodiffeq_b200.solvers as Sv
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200.conditions import IVP
    from neurodiffeq_b200.generators import Generator1D
    from neurodiffeq_b200.networks import FCNN
    torch.manual_seed(0)
    net = FCNN(1, 1, hidden_units=(16, 16))
    solver = Sv.Solver1D(lambda u, t: [diff(u, t) + u], [IVP(0.0, 1.0)], t_min=0.0, t_max=2.0, nets=[net],
                         train_generator=Generator1D(48, 0.0, 2.0, "equally-spaced-noisy"),
                         valid_generator=Generator1D(32, 0.0, 2.0, "equally-spaced"), n_batches_valid=1,
                         optimizer=torch.optim.Adam(net.parameters(), lr=5e-3))
    solver.fit(400, tqdm_file=None)
    ts = np.linspace(0.0, 2.0, 41)
    u = solver.get_solution(best=True)(ts, to_numpy=True)
    assert solver.metrics_history["valid_loss"][-1] < 1e-3 * solver.metrics_history["valid_loss"][0]
    assert np.abs(u - np.exp(-ts)).max() < 2e-2
    assert solver.lowest_loss == min(solver.metrics_history["valid_loss"])


def test_bundle_and_spherical_solutions_cpu():
    """BundleSolution1D takes (t, *bundle parameters); SolutionSpherical takes (r, theta, phi); both reshape to the
    first argument's shape and honour the conditions (reference solvers.py:977-1013, 1184-1187)."""
    wl, solver, nets, coords_np = make_solver("c5", 64)
    solver.fit(1, tqdm_file=None)
    sol = solver.get_solution(best=False)
    t0 = np.zeros((3, 4))
    zeta, omega = np.full((3, 4), 0.2), np.full((3, 4), 1.0)
    u0, v0 = np.linspace(-1, 1, 12).reshape(3, 4), np.linspace(1, -1, 12).reshape(3, 4)
    u, v = sol(t0, zeta, omega, u0, v0, to_numpy=True)
    assert u.shape == (3, 4) and v.shape == (3, 4)
    np.testing.assert_allclose(u, u0, atol=1e-7)                      # BundleIVP: u(t0) = u_0 taken per point
    np.testing.assert_allclose(v, v0, atol=1e-7)
    wl, solver, nets, coords_np = make_solver("c4", 48)
    solver.fit(1, tqdm_file=None)
    r = np.full(5, 0.1)
    th, ph = np.linspace(0.3, 2.5, 5), np.linspace(0.1, 6.0, 5)
    inner = solver.get_solution(best=True)(r, th, ph, to_numpy=True)
    assert inner.shape == (5,) and np.allclose(inner, inner[0])       # constant Dirichlet value on the inner shell
    res = solver.get_residuals(r + 1.0, th, ph, to_numpy=True)
    assert res.shape == (5,) and np.isfinite(res).all()
    internals = solver.get_internals(["nets", "conditions"], return_type="dict")
    assert set(internals) == {"nets", "conditions"}


def test_compute_func_val_hook_and_spherical_enforcer_cpu():
    """Solver hooks of the reference API (solvers.py:267-279, 894-916): an overridden ``compute_func_val`` and a spherical
    ``enforcer`` decide how a condition is applied; the fused solvers trace through them."""
    import neurodiffeq_b200.solvers as Sv
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200.conditions import IVP, NoCondition
    from neurodiffeq_b200.generators import Generator1D, GeneratorSpherical
    from neurodiffeq_b200.networks import FCNN

    class Shifted(Sv.Solver1D):
        def compute_func_val(self, net, cond, *coordinates):
            return cond.enforce(net, *coordinates) + 2.0            # u = parameterised value + 2

    g = Generator1D(16, 0.0, 1.0, "equally-spaced")
    s = Shifted(lambda u, t: [diff(u, t)], [IVP(0.0, 1.0)], nets=[FCNN(1, 1, hidden_units=(8,))], train_generator=g,
                valid_generator=g)
    assert abs(float(s.get_solution(best=False)(torch.zeros(1))) - 3.0) < 1e-12

    seen = []

    def enforcer(net, cond, coordinates):
        seen.append(len(coordinates))
        r, th, ph = coordinates
        return cond.enforce(net, r, th, ph) * (r - 1.0)             # vanishes on the unit sphere

    gs = GeneratorSpherical(24, 0.5, 1.5)
    sp = Sv.SolverSpherical(lambda u, r, th, ph: [diff(u, r)], [NoCondition()], nets=[FCNN(3, 1, hidden_units=(8,))],
                            train_generator=gs, valid_generator=gs, enforcer=enforcer)
    assert seen and seen[0] == 3
    sp.fit(1, tqdm_file=None)
    u = sp.get_solution(best=False)(np.ones(4), np.linspace(0.3, 2.0, 4), np.linspace(0.1, 3.0, 4), to_numpy=True)
    assert np.allclose(u, 0.0, atol=1e-12)
