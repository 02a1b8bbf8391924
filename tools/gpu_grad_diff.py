"""Diagnostic (GPU box): per-parameter-tensor gradient error of a workload against its golden vectors.
usage: PINNJET_TC=2 python tools/gpu_grad_diff.py c2 [c5 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import workloads  # noqa: E402
from helpers import build_fused, product_namespace
from conftest import load_golden  # noqa: E402
from test_kernels_gpu import run_fused  # noqa: E402

for key in sys.argv[1:] or ["c2"]:
    wl0 = workloads.build(product_namespace(), key)
    gold = load_golden(wl0.name)
    wl, nets, conds, fp = build_fused(key, params=gold["params"])
    u, r, loss_eval, r2, loss_train, grads = run_fused(fp, gold["coords"])
    print("==", key, "PINNJET_TC =", os.environ.get("PINNJET_TC"))
    for i, (g, ref) in enumerate(zip(grads, gold["grads"])):
        g, ref = np.asarray(g, dtype=np.float64), np.asarray(ref, dtype=np.float64)
        err = np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-300)
        ratio = (g.reshape(-1)[:4] / np.where(ref.reshape(-1)[:4] == 0, 1, ref.reshape(-1)[:4]))
        print(f"  tensor {i} shape {g.shape}: rel err {err:.3e}  first ratios {np.round(ratio, 4)}")
