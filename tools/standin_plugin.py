"""pytest plugin: dry-run the GPU test files against the CPU stand-in engine (tests/cpu_engine.py) before spending GPU time.

    PYTHONPATH=tools python -m pytest -p standin_plugin tests/test_solvers_gpu.py tests/test_losses_gpu.py \\
        tests/test_kernels_gpu.py tests/test_conditions_gpu.py -m gpu -q

The host-side flow of every test (tracing, solver loop, shapes, tolerances against the oracle) runs for real; the CUDA engine
is replaced by the float64 numpy mirror, `.cuda()` / `torch.cuda.synchronize()` become no-ops.  Tests that assert on the CUDA
library itself or create tensors with device="cuda" fail here by design.  Test infrastructure only."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


@pytest.fixture(autouse=True)
def _stand_in_engine(monkeypatch):
    import neurodiffeq_b200.engine as E
    import neurodiffeq_b200.solvers as S
    from cpu_engine import CpuFusedProblem
    monkeypatch.setattr(S, "FusedProblem", CpuFusedProblem)
    monkeypatch.setattr(E, "FusedProblem", CpuFusedProblem)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self.double() if self.is_floating_point() else self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)
