"""Every operator of the reference's ``operators.py`` (:15-432) against golden vectors produced by the unmodified
reference (tests/golden/generate_operators.py): once on eager tensors (autograd, what user post-processing code gets)
and once on traced symbols through the residual program (what the fused kernels evaluate); plus the vector-calculus
identities of reference tests/test_operators_identities.py:57-143 in all three coordinate systems."""
import os

import numpy as np
import pytest
import torch

import workloads
from conftest import GOLDEN_DIR
from cpu_engine import CpuFusedProblem
from neurodiffeq_b200 import operators as ops
from neurodiffeq_b200.conditions import NoCondition
from neurodiffeq_b200.networks import FCNN

GOLD = np.load(os.path.join(GOLDEN_DIR, "operators_n48.npz"))


@pytest.mark.parametrize("name", workloads.OPERATOR_NAMES)
def test_operator_on_tensors_matches_reference(name):
    c = [torch.tensor(v, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for v in GOLD["coords"]]
    res = getattr(ops, name)(*workloads.operator_arguments(name, c))
    res = res if isinstance(res, (tuple, list)) else (res,)
    got = np.stack([r.detach().numpy()[:, 0] for r in res])
    np.testing.assert_allclose(got, GOLD[name], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", workloads.OPERATOR_NAMES)
def test_operator_on_traced_symbols_matches_reference(name):
    def eqs(u, a, b, d):
        res = getattr(ops, name)(*workloads.operator_arguments(name, (a, b, d)))
        res = res if isinstance(res, (tuple, list)) else (res,)
        return [r + 0 * u for r in res]

    fp = CpuFusedProblem([FCNN(3, 1, hidden_units=(4,))], [NoCondition()], eqs, 3)
    _, r, _ = fp.forward([torch.tensor(v, dtype=torch.float64) for v in GOLD["coords"]])
    # the lowered program carries float32 immediates for non-representable constants -> 1e-6 relative
    scale = 1.0 + np.abs(GOLD[name])
    assert np.max(np.abs(r.numpy() - GOLD[name]) / scale) < 2e-6


def _coords(seed=0):
    rs = np.random.RandomState(seed)
    return [torch.tensor(v, dtype=torch.float64).reshape(-1, 1).requires_grad_(True)
            for v in (0.5 + rs.rand(20), 0.4 + 2.0 * rs.rand(20), 0.3 + 1.7 * rs.rand(20))]


@pytest.mark.parametrize("system", ["cartesian", "spherical", "cylindrical"])
def test_vector_calculus_identities_in_every_coordinate_system(system):
    c = _coords()
    f = workloads.operator_fields(*c)
    G, D, K, L, VL = {
        "cartesian": (ops.grad, ops.div, ops.curl, ops.laplacian, ops.vector_laplacian),
        "spherical": (ops.spherical_grad, ops.spherical_div, ops.spherical_curl, ops.spherical_laplacian,
                      ops.spherical_vector_laplacian),
        "cylindrical": (ops.cylindrical_grad, ops.cylindrical_div, ops.cylindrical_curl, ops.cylindrical_laplacian,
                        ops.cylindrical_vector_laplacian),
    }[system]
    zero = lambda t: float(t.detach().abs().max()) < 1e-9                               # noqa: E731
    g = G(f[0], *c)
    assert zero(D(*g, *c) - L(f[0], *c))                                       # div grad = laplacian
    assert all(zero(k) for k in K(*g, *c))                                     # curl grad = 0
    assert zero(D(*K(*f, *c), *c))                                             # div curl = 0
    gd = G(D(*f, *c), *c)
    cc = K(*K(*f, *c), *c)
    for lhs, a, b in zip(VL(*f, *c), gd, cc):                                  # vector laplacian = grad div - curl curl
        assert zero(lhs - (a - b))


def test_coordinate_transformations_round_trip():
    c = _coords(1)
    r, th, ph = c[0], 0.2 + c[1], c[2]
    back = ops.cartesian_to_spherical(*ops.spherical_to_cartesian(r, th, ph))
    for a, b in zip(back, (r, th, ph)):
        assert torch.allclose(a, b, rtol=1e-12)
    back = ops.cartesian_to_cylindrical(*ops.cylindrical_to_cartesian(r, ph, c[1]))
    for a, b in zip(back, (r, ph, c[1])):
        assert torch.allclose(a, b, rtol=1e-12)
