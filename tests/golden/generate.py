"""Generate golden vectors by running the UNMODIFIED reference (imported in place from /root/reference).

Run in the build container only (``python tests/golden/generate.py``); the GPU box has no /root/reference, so the
``.npz`` files written next to this script are committed and are what the tests read.

For every workload of workloads.py (BASELINE.json's C1..C5, the extension cases x1.. and the fallback cases y1..) at N points the script evaluates exactly the closure of
the reference (solvers.py:369-395): ``funcs = cond.enforce(net, *coords)``; ``r = cat(diff_eqs(*funcs, *coords))``;
``loss = (r**2).mean()``; ``loss.backward()`` -- in float64 (the reference's import default) on inputs whose
values are float32-representable (so that the fp64 result is "the exact answer" for the fp32 inputs the CUDA path
sees), plus once more in float32 to record the reference's own fp32-vs-fp64 noise (the parity floor).

Stored per workload: coords [d0,N] f32; params (state_dict order, per distinct net) f32; u [n_funcs,N] f64;
residual [n_eq,N] f64; loss f64; grads (same order as params) f64; residual32/loss32/grad32 from the fp32 re-run.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402
import workloads  # noqa: E402

GOLDEN_N = {"c1": 256, "c2": 256, "c3": 256, "c4": 256, "c5": 256, "x1": 256, "x2": 256, "x3": 256, "x4": 256, "x5": 256,
            "x6": 256, "x7": 256, "x8": 256, "x9": 256, "y1": 256, "y2": 256, "y3": 256}


def reference_namespace():
    ref_shim.import_reference()
    from neurodiffeq import diff
    from neurodiffeq.networks import FCNN, SinActv, Resnet
    from neurodiffeq.conditions import (IVP, BundleIVP, DirichletBVP2D, IBVP1D, DirichletBVPSpherical, NoCondition,
                                        DoubleEndedBVP1D, EnsembleCondition)
    from neurodiffeq.operators import spherical_laplacian, laplacian, grad, div, curl
    return types.SimpleNamespace(
        diff=diff, FCNN=FCNN, Resnet=Resnet, SinActv=SinActv, IVP=IVP, BundleIVP=BundleIVP, DirichletBVP2D=DirichletBVP2D,
        IBVP1D=IBVP1D, DirichletBVPSpherical=DirichletBVPSpherical, NoCondition=NoCondition,
        DoubleEndedBVP1D=DoubleEndedBVP1D, EnsembleCondition=EnsembleCondition, spherical_laplacian=spherical_laplacian, laplacian=laplacian, grad=grad, div=div, curl=curl)


def distinct(nets):
    seen, out = set(), []
    for n in nets:
        if id(n) not in seen:
            seen.add(id(n))
            out.append(n)
    return out


def run_closure(wl, nets, conds, coords_np, dtype):
    """The reference closure (solvers.py:369-395) on fixed points; returns numpy results."""
    for n in distinct(nets):
        n.to(dtype)
        for p in n.parameters():
            p.grad = None
    coords = [torch.tensor(c, dtype=dtype).reshape(-1, 1).requires_grad_(True) for c in coords_np]
    funcs = [c.enforce(n, *coords) for n, c in zip(nets, conds)]
    eqs = workloads.bundle_eq_wrapper(wl)
    residuals = torch.cat(eqs(*funcs, *coords), dim=1)
    loss = (residuals ** 2).mean()
    loss.backward()
    grads = [p.grad.detach().cpu().numpy().copy() for n in distinct(nets) for p in n.parameters()]
    return (np.concatenate([f.detach().numpy().T for f in funcs]), residuals.detach().numpy().T.copy(),   # (N, k) blocks -> k rows
            float(loss.item()), grads)


def main():
    nd = reference_namespace()
    only = sys.argv[1:]   # e.g. `generate.py x1 x2`: (re)generate just these; default: every workload
    for key in workloads.NAMES + workloads.EXTRA_NAMES + workloads.FALLBACK_NAMES:
        if only and key not in only:
            continue
        wl = workloads.build(nd, key)
        n_pts = GOLDEN_N[key]
        torch.manual_seed(0)
        nets = wl.make_nets()
        conds = wl.make_conditions()
        for n in distinct(nets):  # make parameter values float32-representable
            for p in n.parameters():
                p.data = p.data.float().double()
        coords = workloads.sample_coords(wl, n_pts, seed=1234)
        params = [p.detach().numpy().astype(np.float32) for n in distinct(nets) for p in n.parameters()]
        u64, r64, loss64, g64 = run_closure(wl, nets, conds, coords, torch.float64)
        u32, r32, loss32, g32 = run_closure(wl, nets, conds, coords, torch.float32)
        for n in distinct(nets):
            n.to(torch.float64)
        out = dict(coords=coords, u=u64, residual=r64, loss=np.float64(loss64),
                   residual32=r32.astype(np.float32), loss32=np.float32(loss32), n_params=np.int64(len(params)))
        for i, (p, g, g_32) in enumerate(zip(params, g64, g32)):
            out[f"param_{i}"] = p
            out[f"grad_{i}"] = g
            out[f"grad32_{i}"] = g_32.astype(np.float32)
        path = os.path.join(HERE, f"{wl.name}_n{n_pts}.npz")
        np.savez_compressed(path, **out)
        gnorm = np.sqrt(sum((g ** 2).sum() for g in g64))
        dnorm = np.sqrt(sum(((g - h) ** 2).sum() for g, h in zip(g64, g32)))
        print(f"{wl.name}: N={n_pts} loss={loss64:.9e} rms(r)={np.sqrt((r64 ** 2).mean()):.4e} "
              f"fp32 self-noise: max|dr|/rms={np.abs(r64 - r32).max() / np.sqrt((r64 ** 2).mean()):.2e} "
              f"dloss={abs(loss64 - loss32) / loss64:.2e} dgrad={dnorm / gnorm:.2e} -> {os.path.basename(path)}")


if __name__ == "__main__":
    main()
