"""The solver API contract of the reference, re-stated from its own tests (reference tests/test_solvers.py:42-290) for the
fused solvers on the CPU stand-in engine: legacy names and their warnings, error types, history / batch helpers, tqdm
routing, residual shapes, default generators, internal variables, best-network bookkeeping."""
import random
import sys

import numpy as np
import pytest
import torch

from cpu_engine import CpuFusedProblem
from neurodiffeq_b200 import diff
from neurodiffeq_b200.conditions import IVP
from neurodiffeq_b200.generators import Generator1D
from neurodiffeq_b200.networks import FCNN
from neurodiffeq_b200.solvers import BaseSolver, GenericSolver, Solver1D, Solver2D, SolverSpherical, BundleSolver1D

T_MIN, T_MAX = 0.0, 1.0
DIFF_EQS = lambda u, t: [diff(u, t) + u]          # noqa: E731
CONDITIONS = [IVP(0, 1)]


@pytest.fixture(autouse=True)
def cpu_engine(monkeypatch):
    import neurodiffeq_b200.solvers as S
    monkeypatch.setattr(S, "FusedProblem", CpuFusedProblem)


@pytest.fixture
def generators():
    return dict(train=Generator1D(64, t_min=T_MIN, t_max=T_MAX, method="uniform"),
                valid=Generator1D(64, t_min=T_MIN, t_max=T_MAX, method="equally-spaced"))


def make(generators, **kw):
    args = dict(diff_eqs=DIFF_EQS, conditions=CONDITIONS, train_generator=generators["train"],
                valid_generator=generators["valid"], n_input_units=1, n_output_units=1)
    args.update(kw)
    return GenericSolver(**args)


@pytest.fixture
def solver(generators):
    return make(generators)


def test_legacies(solver, generators):
    solver.fit(1, tqdm_file=None)
    assert solver.batch == solver._batch
    with pytest.warns(FutureWarning):
        assert solver._batch_examples == solver._batch
    with pytest.raises(TypeError), pytest.warns(FutureWarning):      # (residuals, zeros) criteria are gone since v0.4.0
        make(generators, criterion=lambda residuals, zeros: (residuals ** 2).mean()).fit(1, tqdm_file=None)

    class SolverWithLegacyAdditionalLoss(GenericSolver):
        def additional_loss(self, funcs, key):
            return 0

    with pytest.raises(TypeError), pytest.warns(FutureWarning):
        SolverWithLegacyAdditionalLoss(diff_eqs=DIFF_EQS, conditions=CONDITIONS, train_generator=generators["train"],
                                       valid_generator=generators["valid"], n_input_units=1, n_output_units=1).fit(1, tqdm_file=None)
    with pytest.warns(FutureWarning):
        make(generators, shuffle=True)
    with pytest.warns(FutureWarning):
        make(generators, criterion="l2")


def test_missing_generator(generators):
    for kw in (dict(valid_generator=None), dict(train_generator=None), dict(train_generator=None, valid_generator=None)):
        with pytest.raises(ValueError):
            make(generators, **kw)


def test_history_and_batch_helpers(solver):
    for key in ("train", "valid"):
        with pytest.raises(KeyError):
            solver._update_history(1.0, metric_type="bad name", key=key)
        for _ in range(3):
            r = random.random()
            getattr(solver, f"_update_{key}_history")(value=r, metric_type="loss")
            assert solver.metrics_history[f"{key}_loss"][-1] == r
        batch = getattr(solver, f"_generate_{key}_batch")()
        assert all(torch.equal(a, b) for a, b in zip(batch, solver._batch[key]))


def test_no_validation_lbfgs_and_early_stopping(solver, generators):
    solver.n_batches["valid"] = 0
    solver.fit(1, tqdm_file=None)
    nets = [FCNN()]
    make(generators, nets=nets, optimizer=torch.optim.LBFGS(params=nets[0].parameters(), lr=1e-3)).fit(1, tqdm_file=None)

    def stop(s):
        s._stop_training = True

    s2 = make(generators)
    s2.fit(max_epochs=10, callbacks=[stop], tqdm_file=None)
    assert s2.global_epoch == 1


def test_invalid_get_internals(solver):
    with pytest.raises(ValueError):
        solver.get_internals(["generator"], return_type="bad type")


def test_tqdm(solver, capfd):
    desc = "Training Progress"
    solver.fit(max_epochs=3, tqdm_file=sys.stdout)
    out, err = capfd.readouterr()
    assert desc in out and desc not in err
    solver.fit(max_epochs=3, tqdm_file=sys.stderr)
    out, err = capfd.readouterr()
    assert desc not in out and desc in err
    solver.fit(max_epochs=3, tqdm_file=None)
    out, err = capfd.readouterr()
    assert desc not in out and desc not in err


@pytest.mark.parametrize("best", [True, False])
@pytest.mark.parametrize("ts", [np.linspace(0, 1, 10), torch.linspace(0, 1, 10)])
@pytest.mark.parametrize("to_numpy", [True, False])
@pytest.mark.parametrize("first_shape", [(-1,), (-1, 1)])
def test_get_residual(solver, best, ts, to_numpy, first_shape):
    solver.fit(1, tqdm_file=None)
    ts = ts.reshape(*first_shape)
    rs = solver.get_residuals(ts, to_numpy=to_numpy, best=best)
    assert isinstance(rs, np.ndarray if to_numpy else torch.Tensor)
    assert rs.shape == rs.reshape(first_shape).shape


def test_generic_solution(solver):
    solution = solver.get_solution(best=False)
    assert (solution(torch.zeros((1, 1))) == 1).all()


@pytest.mark.parametrize("SolverClass", [Solver1D, Solver2D, SolverSpherical, BundleSolver1D])
def test_missing_domain(SolverClass, generators):
    with pytest.raises(ValueError):
        SolverClass(DIFF_EQS, CONDITIONS)
    with pytest.raises(ValueError):
        SolverClass(DIFF_EQS, CONDITIONS, train_generator=generators["train"])
    with pytest.raises(ValueError):
        SolverClass(DIFF_EQS, CONDITIONS, valid_generator=generators["valid"])


def test_default_generator():
    """Domain bounds instead of generators build the reference's default generators (solvers.py:1107-1160, 1519-1570,
    854-890).  The fused solvers trace the problem at construction, so the definitions must be consistent (the reference
    would only notice at the first batch)."""
    from neurodiffeq_b200.conditions import NoCondition, DirichletBVPSpherical
    s1 = Solver1D(DIFF_EQS, CONDITIONS, t_min=0, t_max=1)
    assert s1.generator["train"].size == 32
    s2 = Solver2D(lambda u, x, y: [diff(u, x) + diff(u, y)], [NoCondition()], xy_min=(0, 0), xy_max=(1, 1))
    assert s2.generator["train"].size == 32 * 32
    cond = DirichletBVPSpherical(0.1, lambda th, ph: 0 * th, 1.0, lambda th, ph: 0 * th + 1)
    s3 = SolverSpherical(lambda u, r, th, ph: [diff(u, r)], [cond], r_min=0.1, r_max=1)
    assert len(s3.generator["train"].get_examples()) == 3


@pytest.mark.parametrize("SolverClass", [Solver1D, BundleSolver1D])
def test_get_internals_variables(SolverClass, generators):
    s = SolverClass(DIFF_EQS, CONDITIONS, train_generator=generators["train"], valid_generator=generators["valid"])
    d1 = BaseSolver._get_internal_variables(s)
    d2 = s._get_internal_variables()
    for k in d1:
        assert k in d2, f"{k} not in {d2.keys()}"


def test_best_nets_with_training(generators):
    s = make(generators, n_batches_valid=0)
    assert s.best_nets is None and s.lowest_loss is None
    s.fit(1, tqdm_file=None)
    assert s.best_nets is not None and s.lowest_loss is not None
    nets = [FCNN()]
    with pytest.warns(RuntimeWarning):
        make(generators, nets=nets, optimizer=torch.optim.LBFGS(nets[0].parameters(), lr=1e-3), n_batches_valid=0)
