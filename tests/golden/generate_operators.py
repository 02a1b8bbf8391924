"""Golden vectors for ``operators.py`` from the UNMODIFIED reference: every operator of neurodiffeq/operators.py:15-432 on
closed-form fields at 48 random points (float64).  ``python tests/golden/generate_operators.py`` (build container only)
-> ``operators_n48.npz``: the coordinates and, per operator, the reference outputs.  The fields are defined in
``workloads.operator_fields`` so that the tests evaluate the same closed forms on tensors and on traced symbols."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402
import workloads  # noqa: E402


def main():
    ref_shim.import_reference()
    import neurodiffeq.operators as R
    torch.set_default_dtype(torch.float64)
    rs = np.random.RandomState(7)
    coords = np.stack([0.5 + rs.rand(48), 0.4 + 2.0 * rs.rand(48), 0.3 + 1.7 * rs.rand(48)])
    out = dict(coords=coords)
    for name in workloads.OPERATOR_NAMES:
        c = [torch.tensor(v).reshape(-1, 1).requires_grad_(True) for v in coords]
        res = getattr(R, name)(*workloads.operator_arguments(name, c))
        res = res if isinstance(res, (tuple, list)) else (res,)
        out[name] = np.stack([r.detach().numpy()[:, 0] for r in res])
        print(f"{name:30s} {out[name].shape}")
    np.savez_compressed(os.path.join(HERE, "operators_n48.npz"), **out)


if __name__ == "__main__":
    main()
