"""GPU-side stage-by-stage comparison of the kernels' buffers with the numpy mirror (oracle/jet_numpy.py).

Usage (on the GPU box):  python tests/dev/gpu_debug.py c2 [N]
Prints, per stage, the max abs error: u, residual, z-jets of every hidden layer (from the workspace), seeds,
gradient of every parameter tensor.  Diagnostic tool; not part of the product.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workloads  # noqa: E402
from helpers import build_fused, get_params  # noqa: E402
from oracle import jet_numpy  # noqa: E402


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    wl, nets, conds, fp = build_fused(key, seed=1)
    tp = fp.tp
    coords_np = workloads.sample_coords(wl, n, seed=4)
    per_net = []
    for nd in tp.nets:
        per_net.append([p.detach().cpu().numpy().astype(np.float64) for p in nd.parameters()])
    ref = jet_numpy.run_traced(tp, per_net, coords_np)
    info = fp.plan_info(n)
    print({k: v for k, v in info.items() if k not in ("hp", "zj_off")})
    coords = [torch.from_numpy(c).cuda() for c in coords_np]
    u, r, ss = fp.forward(coords, want_sumsq=True)
    torch.cuda.synchronize()
    print("eval  u   max err", np.abs(u.cpu().numpy() - ref["u"]).max())
    print("eval  r   max err", np.abs(r.cpu().numpy() - ref["residual"]).max(), "rms", np.sqrt((ref["residual"] ** 2).mean()))
    print("eval  loss", float(ss) / (n * fp.n_eq), "ref", ref["loss"])
    fp.grad.zero_()
    s2, r2 = fp.residual_grad(coords, want_residual=True)
    torch.cuda.synchronize()
    print("train r   max err", np.abs(r2.cpu().numpy() - ref["residual"]).max())
    print("train loss", float(s2) / (n * fp.n_eq))
    ws = fp.workspace.cpu().numpy()
    T, RS, C = info["T"], info["RS"], info["C"]
    n_tiles = info["n_tiles"]
    zj = ws[info["ws_zj"]:info["ws_zj"] + 4 * info["zj_tile_floats"] * n_tiles].view(np.float32).reshape(n_tiles, -1)
    seeds = ws[info["ws_seed"]:info["ws_seed"] + 4 * tp.n_yrows * T * n_tiles].view(np.float32).reshape(n_tiles, tp.n_yrows, T)
    for k, nd in enumerate(tp.nets):
        for h in range(1, len(nd.linears)):
            zref = ref["z_store"][k][h - 1].copy()  # [C, width, N]
            if nd.act == 0:
                zref[0] = np.tanh(zref[0])  # K1 stores tanh(z0) in channel 0 for tanh nets
            hp = info["hp"][k][h]
            blk = zj[:, info["zj_off"][k][h]: info["zj_off"][k][h] + hp * RS].reshape(n_tiles, hp, RS)[:, :, :C * T]
            blk = blk.reshape(n_tiles, hp, C, T).transpose(2, 1, 0, 3).reshape(C, hp, n_tiles * T)[:, :nd.widths[h], :n]
            print(f"net{k} hidden{h} z-jets max err", np.abs(blk - zref).max(), "scale", np.abs(zref).max())
    sd = seeds.transpose(1, 0, 2).reshape(tp.n_yrows, n_tiles * T)[:, :n]
    print("seeds max err", np.abs(sd - ref["seeds"]).max(), "scale", np.abs(ref["seeds"]).max())
    got = fp.grads_as_list()
    for i, (g, h) in enumerate(zip(got, ref["grads"])):
        print(f"grad[{i}] shape {g.shape} max err {np.abs(g - h.reshape(g.shape)).max():.3e} scale {np.abs(h).max():.3e}")


if __name__ == "__main__":
    main()
