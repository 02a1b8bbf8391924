"""The CPU oracle (oracle/reference_port.py) against golden vectors made by the unmodified reference."""
import numpy as np
import pytest
import torch

import workloads
from conftest import load_golden
from oracle import reference_port as oracle


@pytest.mark.parametrize("key", workloads.NAMES + workloads.EXTRA_NAMES + workloads.FALLBACK_NAMES)
def test_oracle_matches_reference_golden(key):
    wl = workloads.build(oracle.NAMESPACE, key)
    gold = load_golden(wl.name)
    nets, conds = wl.make_nets(), wl.make_conditions()
    oracle.load_params(nets, gold["params"])
    out = oracle.evaluate(nets, conds, workloads.bundle_eq_wrapper(wl), gold["coords"], dtype=torch.float64)
    # same algorithm, same fp64 arithmetic, same torch: agreement is at rounding level
    np.testing.assert_allclose(out["u"], gold["u"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["residual"], gold["residual"], rtol=1e-10, atol=1e-11)
    assert abs(out["loss"] - gold["loss"]) <= 1e-12 * abs(gold["loss"])
    assert len(out["grads"]) == len(gold["grads"])
    for g, h in zip(out["grads"], gold["grads"]):
        np.testing.assert_allclose(g, h, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(h).max()))


def test_diff_semantics():
    """reference tests/test_neurodiffeq.py:21-106: shape rules, d/dt t^2, unused -> zeros, higher orders."""
    t = torch.linspace(0, 1, 11, dtype=torch.float64).reshape(-1, 1).requires_grad_(True)
    u = t ** 2
    assert torch.allclose(oracle.diff(u, t), 2 * t)
    assert torch.allclose(oracle.diff(u, t, order=2), torch.full_like(t, 2.0))
    assert torch.allclose(oracle.diff(u, t, order=3), torch.zeros_like(t))
    s = torch.rand(11, 1, dtype=torch.float64, requires_grad=True)
    assert torch.equal(oracle.diff(u, s), torch.zeros_like(s))
    with pytest.raises(ValueError):
        oracle.diff(u.reshape(-1), t)
    with pytest.raises(ValueError):
        oracle.diff(u, t[:5])
    e = torch.exp(t)
    for k in range(1, 6):
        assert torch.allclose(oracle.diff(e, t, order=k), e)
