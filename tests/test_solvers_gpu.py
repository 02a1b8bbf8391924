"""The reference-facing API on the GPU: Solver*.fit() drives the fused kernels; a few optimizer steps must track the
CPU oracle (autograd closure + torch Adam in float64) started from the same parameters on the same fixed batch."""
import numpy as np
import pytest
import torch

import workloads
from helpers import product_namespace, get_params, set_params

pytestmark = pytest.mark.gpu


def oracle_training(key, params, coords_np, epochs, lr=1e-3):
    from oracle import reference_port as oracle
    wl = workloads.build(oracle.NAMESPACE, key)
    nets, conds = wl.make_nets(), wl.make_conditions()
    oracle.load_params(nets, params, dtype=torch.float64)
    mods = oracle.distinct_modules(nets)
    opt = torch.optim.Adam([p for m in mods for p in m.parameters()], lr=lr)
    eqs = workloads.bundle_eq_wrapper(wl)
    losses = []
    for _ in range(epochs):
        opt.zero_grad()
        coords = [torch.as_tensor(c, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for c in coords_np]
        _, _, loss = oracle.closure(nets, conds, eqs, coords)
        losses.append(float(loss.detach()))
        opt.step()
    return losses, [p.detach().numpy().copy() for m in mods for p in m.parameters()]


def make_solver(key, n, seed=0, **kw):
    from neurodiffeq_b200 import solvers as S
    from neurodiffeq_b200.generators import PredefinedGenerator
    nd = product_namespace()
    wl = workloads.build(nd, key)
    torch.manual_seed(seed)
    nets, conds = wl.make_nets(), wl.make_conditions()
    coords_np = workloads.sample_coords(wl, n, seed=21)
    gen = PredefinedGenerator(*[c for c in coords_np])
    cls = getattr(S, wl.solver)
    common = dict(nets=nets, train_generator=gen, valid_generator=gen, n_batches_valid=1, **kw)
    if wl.solver == "Solver1D":
        solver = cls(wl.diff_eqs, conds, **common)
    elif wl.solver == "Solver2D":
        solver = cls(wl.diff_eqs, conds, **common)
    elif wl.solver == "SolverSpherical":
        solver = cls(wl.diff_eqs, conds, **common)
    else:
        solver = cls(wl.diff_eqs, conds, eq_param_index=wl.eq_param_index, **common)
    return wl, solver, nets, coords_np


@pytest.mark.parametrize("key", workloads.NAMES)
def test_fit_tracks_oracle_adam(key):
    n, epochs = 1500, 5
    wl, solver, nets, coords_np = make_solver(key, n)
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training(key, params0, coords_np, epochs)
    got = solver.metrics_history["train_loss"]
    assert len(got) == epochs and len(solver.metrics_history["valid_loss"]) == epochs
    np.testing.assert_allclose(got, ref_losses, rtol=2e-4)
    # valid loss of epoch e is evaluated after the step of epoch e = train loss of epoch e+1 (same fixed batch)
    np.testing.assert_allclose(solver.metrics_history["valid_loss"][:-1], ref_losses[1:], rtol=2e-4)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=2e-3, atol=2e-5)
    assert solver.global_epoch == epochs and solver.lowest_loss == min(solver.metrics_history["valid_loss"])


def test_solution_and_residuals_api():
    wl, solver, nets, coords_np = make_solver("c2", 1024)
    solver.fit(2, tqdm_file=None)
    sol = solver.get_solution(best=False)
    xs, ys = np.linspace(0, 1, 17), np.linspace(0, 1, 9)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    u = sol(X, Y, to_numpy=True)
    assert u.shape == X.shape
    # Dirichlet data are met exactly by construction (reference tests/test_conditions.py:352-373)
    np.testing.assert_allclose(u[0, :], np.sin(np.pi * ys), atol=1e-6)
    np.testing.assert_allclose(u[-1, :], 0, atol=1e-6)
    np.testing.assert_allclose(u[:, 0], 0, atol=1e-6)
    np.testing.assert_allclose(u[:, -1], 0, atol=1e-6)
    r = solver.get_residuals(X, Y, to_numpy=True, best=False)
    assert r.shape == X.shape and np.isfinite(r).all()
    rb = solver.get_residuals(torch.tensor(X), torch.tensor(Y), best=True)
    assert isinstance(rb, torch.Tensor) and rb.shape == X.shape
    best = solver.get_solution(best=True)(X, Y, to_numpy=True)
    assert best.shape == X.shape


def test_multi_batch_epoch_accumulates_like_reference():
    """n_batches_train=2 on a fixed batch: gradients add up (no averaging) -> the step equals one with 2x the gradient."""
    wl, s1, nets1, coords = make_solver("c2", 512, n_batches_train=2,
                                        optimizer=None)
    p0 = get_params(nets1)
    s1.run_train_epoch()
    g2 = s1.problem.grad.clone()
    wl, s2, nets2, _ = make_solver("c2", 512)
    set_params(nets2, p0)
    s2.problem.relink()
    s2.run_train_epoch()
    g1 = s2.problem.grad.clone()
    assert torch.allclose(g2, 2 * g1, rtol=1e-4, atol=1e-7)
    assert s1.metrics_history["train_loss"][0] == pytest.approx(s2.metrics_history["train_loss"][0], rel=1e-5)


def test_custom_loss_and_metrics_and_early_stop():
    nd = product_namespace()
    wl = workloads.build(nd, "c2")
    calls = []

    def l1_loss(r, funcs, coords):
        return r.abs().mean()

    def stopper(solver):
        calls.append(solver.local_epoch)
        if solver.local_epoch == 2:
            solver._stop_training = True

    wl, solver, nets, coords_np = make_solver("c2", 700, loss_fn=l1_loss,
                                              metrics={"mean_u": lambda u, x, y: u.mean()})
    solver.fit(10, callbacks=[stopper], tqdm_file=None)
    assert calls == [1, 2] and solver.global_epoch == 2
    assert len(solver.metrics_history["train__mean_u"]) == 2
    # gradient of the custom loss: compare with the oracle's autograd
    from oracle import reference_port as oracle
    owl = workloads.build(oracle.NAMESPACE, "c2")
    onets, oconds = owl.make_nets(), owl.make_conditions()
    wl2, s2, nets2, coords_np = make_solver("c2", 700, loss_fn=l1_loss, seed=5)
    oracle.load_params(onets, get_params(nets2))
    s2.problem.gradbuf.zero_()
    s2._run_epoch("train")
    cs = [torch.as_tensor(c, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for c in coords_np]
    funcs = [c.enforce(n, *cs) for n, c in zip(onets, oconds)]
    res = torch.cat(owl.diff_eqs(*funcs, *cs), 1)
    res.abs().mean().backward()
    ref = np.concatenate([p.grad.numpy().reshape(-1) for p in onets[0].parameters()])
    got = s2.problem.grad.cpu().numpy()
    assert np.linalg.norm(got - ref) <= 2e-4 * np.linalg.norm(ref)
