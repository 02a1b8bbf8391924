"""Named losses of the reference (losses.py:29-35) through the fused solvers: 'l1' and 'infinity' via the custom-loss
path (autograd on the residual matrix -> dL/dr -> kernels), 'h1' as the fused mean square over the augmented residual
rows.  A few Adam steps must track the same training done by autograd on the CPU (oracle, float64)."""
import numpy as np
import pytest
import torch

import workloads
from helpers import get_params
from test_solvers_gpu import make_solver

pytestmark = pytest.mark.gpu


def oracle_training_with_loss(key, params, coords_np, epochs, loss_name, lr=1e-3):
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
        cols = [torch.as_tensor(c, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for c in coords_np]
        funcs = [c.enforce(n, *cols) for n, c in zip(nets, conds)]
        res = torch.cat(eqs(*funcs, *cols), dim=1)
        if loss_name == "l1":                      # losses.py:5-6
            loss = res.abs().mean()
        elif loss_name == "infinity":              # losses.py:13-14
            loss = res.abs().max(dim=1)[0].mean()
        elif loss_name == "h1 semi":               # losses.py:23-26
            loss = (torch.cat(oracle.grad(res, *cols), dim=1) ** 2).mean()
        else:                                      # 'h1', losses.py:17-20
            loss = (torch.cat([res, *oracle.grad(res, *cols)], dim=1) ** 2).mean()
        loss.backward()
        losses.append(float(loss.detach()))
        opt.step()
    return losses, [p.detach().numpy().copy() for m in mods for p in m.parameters()]


@pytest.mark.parametrize("loss_name", ["l1", "infinity", "h1"])
def test_named_losses_track_autograd_training(loss_name):
    key, n, epochs = "c1", 1000, 5
    wl, solver, nets, coords_np = make_solver(key, n, loss_fn=loss_name)
    params0 = get_params(nets)
    solver.fit(epochs, tqdm_file=None)
    ref_losses, ref_params = oracle_training_with_loss(key, params0, coords_np, epochs, loss_name)
    np.testing.assert_allclose(solver.metrics_history["train_loss"], ref_losses, rtol=2e-4)
    for a, b in zip(get_params(nets), ref_params):
        np.testing.assert_allclose(a, b, rtol=2e-3, atol=2e-5)
    if loss_name == "h1":   # the user's residuals, not the augmented rows
        r = solver.get_residuals(torch.linspace(0.5, 2.0, 7), best=False)
        assert isinstance(r, list) and len(r) == wl.n_eq


def test_unknown_and_unsupported_loss_names():
    with pytest.raises(KeyError):
        make_solver("c1", 64, loss_fn="l3")
