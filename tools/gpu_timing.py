"""Diagnostic (GPU box): per-phase cycle counters of CTA 0 from the PJ_TIMING build (libpinnjet_timing.so)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PINNJET_LIB"] = os.path.join(ROOT, "neurodiffeq_b200", "csrc", "libpinnjet_timing.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import workloads  # noqa: E402
from helpers import build_fused  # noqa: E402

K1 = ["setup", "layer0", "gemm", "bar-after-gemm", "epilogue", "output-layer", "program"]
K2 = ["setup", "seeds+z-wait", "last-linear", "adjoint-gemm", "z-wait", "reverse-act", "wgrad", "layer0"]


def main():
    for key in sys.argv[1:] or ["c2"]:
        wl, nets, conds, fp = build_fused(key, seed=0)
        n = wl.default_n
        coords = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=1)]
        for _ in range(3):
            fp.gradbuf.zero_()
            fp.residual_grad(coords)
        torch.cuda.synchronize()
        info = fp.plan_info(n)
        dbg = fp.workspace[640 * 4:(640 + 32) * 4].view(torch.float32).cpu().numpy()
        print(f"== {key}: T={info['T']} grid={info['grid']} n_tiles={info['n_tiles']} "
              f"stages {info['n_stage_fwd']}/{info['n_stage_bwd']} resident {info['resident_fwd']}/{info['resident_bwd']} "
              f"smem {info['smem_fwd']}/{info['smem_bwd']}")
        for name, base, labels in (("K1", 0, K1), ("K2", 16, K2)):
            v = dbg[base:base + len(labels)]
            tot = v.sum()
            print(f"  {name} CTA0 total {tot:.0f} cycles: " + ", ".join(f"{l} {100 * x / tot:.1f}%" for l, x in zip(labels, v)))


if __name__ == "__main__":
    main()
