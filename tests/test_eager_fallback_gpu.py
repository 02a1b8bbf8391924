"""SURVEY.md §8b on the GPU: problems the fused engine refuses run on the autograd path (neurodiffeq_b200/eager.py: torch
autograd on the device) with one warning, and track the fp64 oracle like the fused problems do."""
import numpy as np
import pytest
import torch

import workloads
from helpers import get_params, rel_l2
from test_solvers_gpu import make_solver, oracle_training
from test_losses_gpu import oracle_training_with_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fresh_warning_registry(monkeypatch):
    """the fallback warns ONCE per reason and process: every test starts with a clean slate"""
    import neurodiffeq_b200.eager as E
    monkeypatch.setattr(E, "_WARNED", set())


@pytest.mark.parametrize("key", ["y1", "y2"])      # y3 (2-D, fourth order) is covered on the CPU against the reference goldens
def test_refused_problem_trains_on_the_autograd_path_gpu(key):
    n, epochs = 500, 5
    with pytest.warns(RuntimeWarning, match="falling back to the autograd path"):
        wl, solver, nets, coords_np = make_solver(key, n)
    assert solver.problem.is_eager and solver.problem.device.type == "cuda" and solver.problem.kernel_launches == 0
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training(key, params0, coords_np, epochs)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=2e-4)
    assert rel_l2(get_params(nets), ref_params) <= 1e-4
    t = torch.linspace(0.1, 1.9, 9)
    u, r = solver.get_solution(best=False)(t), solver.get_residuals(t, best=False)
    assert u.shape == (9,) and r.shape == (9,) and u.is_cuda


def test_h1_on_a_second_order_problem_falls_back_gpu():
    key, n, epochs = "x6", 300, 4
    with pytest.warns(RuntimeWarning, match="falling back to the autograd path"):
        wl, solver, nets, coords_np = make_solver(key, n, loss_fn="h1")
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_with_loss(key, params0, coords_np, epochs, "h1")
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=5e-4)
    assert rel_l2(get_params(nets), ref_params) <= 2e-4


def test_fused_problems_do_not_fall_back_gpu():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)
        wl, solver, nets, coords_np = make_solver("c2", 256)
    assert not getattr(solver.problem, "is_eager", False)


def test_live_solution_follows_training_gpu():
    """get_solution(copy=False, best=False) on the FUSED path: the solution's engine re-adopts the live parameters."""
    wl, solver, nets, coords_np = make_solver("c1", 400)
    solver.fit(2, tqdm_file=None)
    live = solver.get_solution(copy=False, best=False)
    t = torch.linspace(0.2, 5.0, 11)
    u1 = [x.clone() for x in live(t)]
    solver.fit(3, tqdm_file=None)
    u2 = live(t)
    fresh = solver.get_solution(copy=True, best=False)(t)
    for a, b, c in zip(u2, fresh, u1):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-6, atol=1e-7)
        assert float((a - c).abs().max()) > 0
    solver.fit(1, tqdm_file=None)
    assert np.isfinite(solver.metrics_history["train_loss"][-1])

