"""Behavioural contract of the solver objects, as the reference's own test-suite pins it (reference
tests/test_solvers.py:42-290), checked for the fused solvers on the CPU stand-in engine.  Grouped by concern:
deprecations, constructor validation, bookkeeping helpers, fit-loop controls, evaluation helpers."""
import random
import sys

import numpy as np
import pytest
import torch

from cpu_engine import CpuFusedProblem
from neurodiffeq_b200 import diff
from neurodiffeq_b200.conditions import IVP, NoCondition, DirichletBVPSpherical
from neurodiffeq_b200.generators import Generator1D
from neurodiffeq_b200.networks import FCNN
from neurodiffeq_b200 import solvers as fused

DECAY = lambda u, t: [diff(u, t) + u]          # noqa: E731   u' = -u, u(0) = 1
START_AT_ONE = [IVP(0, 1)]
BAR = "Training Progress"


@pytest.fixture(autouse=True)
def _stand_in_engine(monkeypatch):
    monkeypatch.setattr(fused, "FusedProblem", CpuFusedProblem)


def points(method):
    return Generator1D(64, t_min=0.0, t_max=1.0, method=method)


def build(cls=fused.GenericSolver, **overrides):
    kw = dict(diff_eqs=DECAY, conditions=START_AT_ONE, train_generator=points("uniform"),
              valid_generator=points("equally-spaced"), n_input_units=1, n_output_units=1)
    kw.update(overrides)
    return cls(**kw)


def quiet_fit(solver, epochs=1, **kw):
    solver.fit(epochs, tqdm_file=None, **kw)
    return solver


class TestDeprecations:
    def test_batch_aliases(self):
        s = quiet_fit(build())
        assert s.batch is s._batch
        with pytest.warns(FutureWarning):
            assert s._batch_examples is s._batch

    def test_two_argument_criterion_is_rejected_with_a_hint(self):
        with pytest.raises(TypeError), pytest.warns(FutureWarning):
            quiet_fit(build(criterion=lambda residuals, zeros: (residuals ** 2).mean()))

    def test_old_additional_loss_signature_is_rejected_with_a_hint(self):
        class OldHook(fused.GenericSolver):
            def additional_loss(self, funcs, key):
                return 0

        with pytest.raises(TypeError), pytest.warns(FutureWarning):
            quiet_fit(build(OldHook))

    @pytest.mark.parametrize("kw", [dict(shuffle=True), dict(criterion="l2"), dict(batch_size=16)])
    def test_retired_keywords_still_construct(self, kw):
        with pytest.warns(FutureWarning):
            build(**kw)


class TestConstructorValidation:
    @pytest.mark.parametrize("missing", [("train_generator",), ("valid_generator",), ("train_generator", "valid_generator")])
    def test_generic_solver_needs_both_generators(self, missing):
        with pytest.raises(ValueError):
            build(**{k: None for k in missing})

    @pytest.mark.parametrize("cls", [fused.Solver1D, fused.Solver2D, fused.SolverSpherical, fused.BundleSolver1D])
    def test_domain_or_generators_are_required(self, cls):
        for kw in ({}, dict(train_generator=points("uniform")), dict(valid_generator=points("uniform"))):
            with pytest.raises(ValueError):
                cls(DECAY, START_AT_ONE, **kw)

    def test_domains_give_the_default_generators(self):
        """(32,) / (32, 32) grids and the spherical sampler (reference solvers.py:1107-1160, 1519-1570, 854-890).  The
        fused solvers trace at construction, so each definition must be consistent with its coordinates."""
        assert fused.Solver1D(DECAY, START_AT_ONE, t_min=0, t_max=1).generator["train"].size == 32
        plane = fused.Solver2D(lambda u, x, y: [diff(u, x) + diff(u, y)], [NoCondition()], xy_min=(0, 0), xy_max=(1, 1))
        assert plane.generator["train"].size == 32 * 32
        shell = DirichletBVPSpherical(0.1, lambda th, ph: 0 * th, 1.0, lambda th, ph: 0 * th + 1)
        ball = fused.SolverSpherical(lambda u, r, th, ph: [diff(u, r)], [shell], r_min=0.1, r_max=1)
        assert len(ball.generator["train"].get_examples()) == 3

    def test_lbfgs_without_validation_warns_about_best_nets(self):
        net = FCNN()
        with pytest.warns(RuntimeWarning):
            build(nets=[net], optimizer=torch.optim.LBFGS(net.parameters(), lr=1e-3), n_batches_valid=0)


class TestBookkeeping:
    @pytest.mark.parametrize("phase", ["train", "valid"])
    def test_history_helpers(self, phase):
        s = build()
        with pytest.raises(KeyError):
            s._update_history(1.0, metric_type="no such metric", key=phase)
        for _ in range(3):
            v = random.random()
            getattr(s, f"_update_{phase}_history")(value=v, metric_type="loss")
            assert s.metrics_history[f"{phase}_loss"][-1] == v

    @pytest.mark.parametrize("phase", ["train", "valid"])
    def test_batch_helpers_return_the_stored_batch(self, phase):
        s = build()
        got = getattr(s, f"_generate_{phase}_batch")()
        assert len(got) == 1 and all(torch.equal(a, b) for a, b in zip(got, s._batch[phase]))

    def test_internal_variables(self):
        s = build()
        with pytest.raises(ValueError):
            s.get_internals(["generator"], return_type="bad type")
        assert set(fused.BaseSolver._get_internal_variables(s)) <= set(s._get_internal_variables())
        for cls in (fused.Solver1D, fused.BundleSolver1D):
            special = cls(DECAY, START_AT_ONE, train_generator=points("uniform"), valid_generator=points("uniform"))
            assert set(fused.BaseSolver._get_internal_variables(special)) <= set(special._get_internal_variables())

    def test_best_nets_appear_with_the_first_epoch(self):
        s = build(n_batches_valid=0)
        assert s.best_nets is None and s.lowest_loss is None
        quiet_fit(s)
        assert s.best_nets is not None and s.lowest_loss is not None


class TestFitLoop:
    def test_validation_can_be_switched_off(self):
        s = build()
        s.n_batches["valid"] = 0
        quiet_fit(s)
        assert len(s.metrics_history["train_loss"]) == 1 and not s.metrics_history["valid_loss"]

    def test_lbfgs_runs(self):
        net = FCNN()
        quiet_fit(build(nets=[net], optimizer=torch.optim.LBFGS(params=net.parameters(), lr=1e-3)))

    def test_a_callback_can_stop_training(self):
        def halt(solver):
            solver._stop_training = True

        assert quiet_fit(build(), epochs=10, callbacks=[halt]).global_epoch == 1

    @pytest.mark.parametrize("stream,shows_out,shows_err", [("stdout", True, False), ("stderr", False, True), (None, False, False)])
    def test_progress_bar_goes_where_it_is_sent(self, capfd, stream, shows_out, shows_err):
        build().fit(max_epochs=3, tqdm_file=getattr(sys, stream) if stream else None)
        out, err = capfd.readouterr()
        assert (BAR in out) == shows_out and (BAR in err) == shows_err


class TestEvaluation:
    @pytest.mark.parametrize("best", [True, False])
    @pytest.mark.parametrize("as_numpy", [True, False])
    @pytest.mark.parametrize("layout", [(-1,), (-1, 1)])
    @pytest.mark.parametrize("container", [np.linspace, torch.linspace])
    def test_residuals_keep_the_shape_of_the_query(self, best, as_numpy, layout, container):
        s = quiet_fit(build())
        ts = container(0, 1, 10).reshape(*layout)
        rs = s.get_residuals(ts, to_numpy=as_numpy, best=best)
        assert isinstance(rs, np.ndarray if as_numpy else torch.Tensor) and tuple(rs.shape) == tuple(ts.shape)

    def test_solution_meets_the_initial_value_exactly(self):
        u = build().get_solution(best=False)
        assert (u(torch.zeros((1, 1))) == 1).all()


def test_legacy_analytic_solutions_become_a_metric():
    """reference solvers.py:151-172: `analytic_solutions` is deprecated but still works, as the metric 'analytic_mse'."""
    with pytest.warns(FutureWarning):
        s = build(analytic_solutions=lambda t: [torch.exp(-t)])
    quiet_fit(s, epochs=2)
    hist = s.metrics_history["train__analytic_mse"]
    assert len(hist) == 2 and all(np.isfinite(hist)) and len(s.metrics_history["valid__analytic_mse"]) == 2
    with pytest.warns(FutureWarning):
        s2 = build(analytic_solutions=lambda t: [torch.exp(-t)], metrics={"analytic_mse": lambda u, t: (u * 0).mean()})
    assert quiet_fit(s2).metrics_history["train__analytic_mse"] == [0.0]
