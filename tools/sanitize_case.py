"""One small residual+gradient evaluation per workload, for compute-sanitizer (memcheck / racecheck / synccheck).
Usage: compute-sanitizer --tool racecheck python tools/sanitize_case.py c2 c5 [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import workloads  # noqa: E402
from helpers import build_fused  # noqa: E402

keys = [a for a in sys.argv[1:] if not a.isdigit()] or ["c2"]
ns = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1500]
for key in keys:
    wl, nets, conds, fp = build_fused(key, seed=0)
    for n in ns:
        coords = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=1)]
        u, r, s = fp.forward(coords, want_sumsq=True)
        fp.gradbuf.zero_()
        s2, _ = fp.residual_grad(coords)
        torch.cuda.synchronize()
        print(key, n, "loss", float(s2) / (n * fp.n_eq), "gradnorm", float(fp.grad.norm()), flush=True)
