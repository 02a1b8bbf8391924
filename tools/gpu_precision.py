"""Diagnostic (GPU box): error statistics of the fused fp32 path vs the fp64 oracle, next to the oracle's own fp32 run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads  # noqa: E402
from helpers import build_fused, get_params, oracle_eval  # noqa: E402


def stats(d):
    return f"max {d.max():.3e} p99.9 {np.percentile(d, 99.9):.3e} p99 {np.percentile(d, 99):.3e} rms {np.sqrt((d**2).mean()):.3e}"


def main():
    for key, n in (("c4", 4097), ("c4", 32768), ("c2", 16384), ("c3", 8192)):
        wl, nets, conds, fp = build_fused(key, seed=7)
        params = get_params(nets)
        coords = workloads.sample_coords(wl, n, seed=99)
        r64 = oracle_eval(key, params, coords, dtype=torch.float64, backward=False)
        r32 = oracle_eval(key, params, coords, dtype=torch.float32, backward=False)
        cd = [torch.from_numpy(c).cuda() for c in coords]
        u, r, _ = fp.forward(cd)
        r = r.cpu().numpy()
        rms = np.sqrt((r64["residual"] ** 2).mean())
        d_ours = np.abs(r - r64["residual"])[0]
        d_ref32 = np.abs(r32["residual"] - r64["residual"])[0]
        i = d_ours.argmax()
        print(f"{key} N={n} rms(r)={rms:.4e}")
        print("   ours   :", stats(d_ours), "worst at", coords[:, i], "ref32 err there", d_ref32[i])
        print("   ref f32:", stats(d_ref32))
        du = np.abs(u.cpu().numpy() - r64["u"])[0]
        du32 = np.abs(r32["u"] - r64["u"])[0]
        print("   u ours :", stats(du), "| ref f32:", stats(du32))


if __name__ == "__main__":
    main()
