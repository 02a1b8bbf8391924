"""Diagnostic (GPU box): event trace of CTA 0 of the tensor-core kernels from the PJ_TIMING build (libpinnjet_timing.so).
usage: PINNJET_TC=2 python tools/gpu_trace_tc.py [workload] [k1|k2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PINNJET_LIB"] = os.path.join(ROOT, "neurodiffeq_b200", "csrc", "libpinnjet_timing.so")
os.environ.setdefault("PINNJET_TC", "2")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import workloads  # noqa: E402
from helpers import build_fused  # noqa: E402


def dump(words, base, cap, title):
    w = words[base:base + cap]
    w = w[w != 0]
    print(f"-- {title}: {len(w)} events")
    prev = 0
    for x in w:
        tag, t = int(x) >> 24, int(x) & 0xFFFFFF
        print(f"   t={t:8d} (+{t - prev:6d})  code={tag & 15} slot={(tag >> 4) & 1} h={tag >> 5}" if title.startswith("K1") else
              f"   t={t:8d} (+{t - prev:6d})  code={tag & 31} h={tag >> 5}")
        prev = t


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "c2"
    which = sys.argv[2] if len(sys.argv) > 2 else "k1"
    wl, nets, conds, fp = build_fused(key, seed=0)
    n = wl.default_n
    coords = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=1)]
    fp.gradbuf.zero_()
    fp.residual_grad(coords)          # allocates the workspace
    import ctypes
    from neurodiffeq_b200 import engine as E
    ptrs, keep = fp._coord_ptrs(coords, n)
    sp = ctypes.byref(fp.spec)
    cs = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        fp.workspace[640 * 4:1024 * 4].zero_()
        E._check(fp.lib.pj_forward_train(sp, fp.prog_train.data_ptr(), len(fp.tp.prog_train), *fp._prog_w_args(), ptrs, n,
                                         fp.pack_buf.data_ptr(), ctypes.c_float(2.0 / (n * fp.n_eq)), None, None, None,
                                         fp.workspace.data_ptr(), fp.workspace.numel(), cs), "k1")
        if which == "k2":
            torch.cuda.synchronize()
            fp.workspace[640 * 4:1024 * 4].zero_()
            E._check(fp.lib.pj_backward(sp, ptrs, n, fp.pack_buf.data_ptr(), fp.grad.data_ptr(), fp.workspace.data_ptr(),
                                        fp.workspace.numel(), cs), "k2")
    torch.cuda.synchronize()
    words = fp.workspace[640 * 4:1024 * 4].view(torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    if which == "k1":
        dump(words, 0, 200, "K1 compute warp 0")
        dump(words, 200, 100, "K1 MMA warp")
        dump(words, 300, 60, "K1 program warp 0")
    else:
        dump(words, 0, 250, "K2 compute warp 0")
        dump(words, 250, 120, "K2 MMA warp")


if __name__ == "__main__":
    main()
