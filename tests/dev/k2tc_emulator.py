"""Thread-level numpy emulation of the EXPERIMENTAL tensor-core reverse kernel (neurodiffeq_b200/csrc/pinnjet_k2tc.cuh).

Not part of the product and not a test of the hardware: the operand encodings (K-major / MN-major reading of the
SWIZZLE_128B images, the M = 64 accumulator lane layout) are facts measured by experiments/tcgen05_probe; here they are taken
as given and the KERNEL'S OWN index algebra and control flow are replayed literally -- per warp and lane: record / seed /
weight addressing in the workspace layout K1 writes, the owner layout, the split-image stores, the TMEM -> staging -> owner
transposition, the accumulator slots, the shared-memory gradient indices and the final read-out -- and the resulting
parameter gradient is compared with the float64 mirror of the algorithm (oracle/jet_numpy.py, itself pinned against the
reference).  Run:  python tests/dev/k2tc_emulator.py [c2 c5 x3:64 ...]   ("x3:64" = workload x3 with 64-wide hidden layers)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workloads  # noqa: E402
from helpers import product_namespace, get_params  # noqa: E402
from neurodiffeq_b200 import symbolic as S  # noqa: E402
from neurodiffeq_b200.engine import pad_scheme, combine_seconds  # noqa: E402
from neurodiffeq_b200.tracing import TracedProblem  # noqa: E402
from oracle import jet_numpy  # noqa: E402

IMG = 128 * 128          # K2T_IMG: bytes of one split image
STAGE_STRIDE = 36


def bf16(x):
    return torch.tensor(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def split3(x):
    x = np.float32(x)
    t1 = bf16(x)
    r = np.float32(x - t1)
    t2 = bf16(r)
    t3 = bf16(np.float32(r - t2))
    return t1, t2, t3


def sw128_off(row, chunk16):
    return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4)


def decode_image(img, split, rows):
    """[rows x 64] matrix of split image `split` as the tensor core sees it (bf16 elements at the swizzled offsets)."""
    out = np.zeros((rows, 64), dtype=np.float64)
    for r in range(rows):
        for u in range(64):
            out[r, u] = img[(split * IMG + sw128_off(r, u >> 3) + (u & 7) * 2) // 2]
    return out


def act_d(act, rec0):
    """(a0, s1, s2, s3) from the record's channel 0 (tanh nets store tanh(z0), sin nets z0)"""
    if act == 0:
        a0 = rec0
        s1 = 1.0 - a0 * a0
        s2 = -2.0 * a0 * s1
        s3 = -2.0 * s1 * s1 - 2.0 * a0 * s2
    else:
        a0, s1 = np.sin(rec0), np.cos(rec0)
        s2, s3 = -a0, -s1
    return a0, s1, s2, s3


def act_from_record(act, z, w, n1, n2, wl):
    a0, s1, s2, _ = act_d(act, z[0])
    a = np.zeros_like(z)
    a[0] = a0
    for f in range(n1):
        a[1 + f] = s1 * z[1 + f]
    if wl:
        q = sum(w[d] * z[1 + d] ** 2 for d in range(wl))
        a[1 + n1] = s2 * q + s1 * z[1 + n1]
    else:
        for s in range(n2):
            a[1 + n1 + s] = s2 * z[1 + s] ** 2 + s1 * z[1 + n1 + s]
    return a


def act_backward(act, z, ab, w, n1, n2, wl):
    """mirror of act_backward<N1,N2,WL> (pinnjet_common.cuh): returns (a, zb)"""
    a0, s1, s2, s3 = act_d(act, z[0])
    C = len(z)
    a, zb = np.zeros(C), np.zeros(C)
    zb0 = s1 * ab[0]
    for f in range(n1):
        zb[1 + f] = s1 * ab[1 + f]
        zb0 += s2 * z[1 + f] * ab[1 + f]
        a[1 + f] = s1 * z[1 + f]
    if wl:
        abL, zL = ab[1 + n1], z[1 + n1]
        q = 0.0
        for d in range(wl):
            wz = w[d] * z[1 + d]
            q += wz * z[1 + d]
            zb[1 + d] += 2.0 * s2 * wz * abL
        zb[1 + n1] = s1 * abL
        zb0 += (s3 * q + s2 * zL) * abL
        a[1 + n1] = s2 * q + s1 * zL
    else:
        for s in range(n2):
            zf, zs, abs_ = z[1 + s], z[1 + n1 + s], ab[1 + n1 + s]
            zb[1 + n1 + s] = s1 * abs_
            zb[1 + s] += 2.0 * s2 * zf * abs_
            zb0 += (s3 * zf * zf + s2 * zs) * abs_
            a[1 + n1 + s] = s2 * zf * zf + s1 * zs
    zb[0] = zb0
    a[0] = a0
    return a, zb


def emulate(key, n_points, T2):
    key, _, width = key.partition(":")      # "x3:64": the workload with all hidden layers `width` wide
    nd_ns = product_namespace()
    wl_ = workloads.build(nd_ns, key)
    torch.manual_seed(2)
    nets, conds = wl_.make_nets(), wl_.make_conditions()
    if width:
        widths = wl_.nets_spec[0][0]
        nets = [nd_ns.FCNN(n_input_units=widths[0], n_output_units=widths[-1], hidden_units=(int(width),) * (len(widths) - 2))
                for _ in nets]
    tp = TracedProblem(nets, conds, workloads.bundle_eq_wrapper(wl_), len(wl_.coord_names), pad_scheme=pad_scheme,
                       combine_seconds=combine_seconds)
    C = tp.n_channels
    n1 = tp.scheme.n1
    n2 = 1 if tp.wl else tp.scheme.n2
    WL = tp.wl
    assert C in (2, 4), f"{key}: {C} channels -- not eligible for the tensor-core kernels"
    for nd in tp.nets:
        assert all(w == 64 for w in nd.widths[1:-1]), "hidden width must be 64"
    params = get_params(nets)
    by_module, it, per_net = {}, iter(params), []
    for nd in tp.nets:
        if id(nd.module) not in by_module:
            by_module[id(nd.module)] = [next(it) for _ in range(len(nd.parameters()))]
        per_net.append(by_module[id(nd.module)])
    coords = workloads.sample_coords(wl_, n_points, seed=4)
    ref = jet_numpy.run_traced(tp, per_net, coords)
    coords_all = tp.extend_coords(coords.astype(np.float64))
    N = n_points

    # ---- workspace exactly as K1 leaves it (records, seeds, weights in K2 record tiles of T2 points) ----
    RS2 = C * T2 + 4
    n_tiles2 = (N + T2 - 1) // T2
    ws_points = n_tiles2 * T2
    zj_off, off = {}, 0
    for n, nd in enumerate(tp.nets):
        for h in range(1, len(nd.linears)):
            zj_off[(n, h)] = off
            off += 64 * RS2
    zj_tile_floats = off
    zj = np.full((n_tiles2, zj_tile_floats), np.nan)
    seeds = np.zeros((n_tiles2, tp.n_yrows * T2))
    n_nets = len(tp.nets)
    wts = np.zeros((n_tiles2, max(1, n_nets * WL) * T2))
    wl_all = S.evaluate_program(tp.prog_w, coords_all, np.zeros((1, N)), n_w=n_nets * WL) if WL else None
    for p in range(N):
        t2, w = p // T2, p % T2
        for n, nd in enumerate(tp.nets):
            for h in range(1, len(nd.linears)):
                z = ref["z_store"][n][h - 1][:, :, p]                       # [C, 64]
                for u in range(64):
                    for c in range(C):
                        v = z[c, u]
                        if c == 0 and nd.act == 0:
                            v = np.tanh(v)
                        zj[t2, zj_off[(n, h)] + u * RS2 + c * T2 + w] = v
        for row in range(tp.n_yrows):
            seeds[t2, row * T2 + w] = ref["seeds"][row, p]
        for r in range(n_nets * WL):
            wts[t2, r * T2 + w] = wl_all[r, p]
    for p in range(N, ws_points):                                           # K1 also writes the padded points of a tile
        t2, w = p // T2, p % T2
        for n, nd in enumerate(tp.nets):
            for h in range(1, len(nd.linears)):
                for u in range(64):
                    for c in range(C):
                        zj[t2, zj_off[(n, h)] + u * RS2 + c * T2 + w] = 0.25

    # ---- parameter offsets (flat theta, torch layout) and the small-gradient slots of the plan ----
    theta_off, o = {}, 0
    for m_id, plist in by_module.items():
        for i, p in enumerate(plist):
            theta_off[(m_id, i)] = o
            o += p.size
    n_theta = o
    g_w0, g_b, g_wl, g_bout, o = {}, {}, {}, {}, 0
    for n, nd in enumerate(tp.nets):
        L = len(nd.linears) - 1
        g_w0[n] = o; o += 64 * nd.widths[0]
        for l in range(L):
            g_b[(n, l)] = o; o += 64
        g_wl[n] = o; o += nd.n_out * 64
        g_bout[n] = o; o += 4
    sgrad = np.zeros(o)
    gpart = np.zeros(n_theta)

    # ---- kernel geometry ----
    T = 128 // C
    PW = 32 // C
    NPP = PW // 2
    NUG = 32 // NPP
    UG = 32 // NUG
    n_sub = T // T2
    assert T % T2 == 0
    rec_sub_floats = 64 * RS2
    n_tiles_tc = (ws_points + T - 1) // T
    n_hh = sum(len(nd.linears) - 2 for nd in tp.nets)
    Dw = np.zeros((n_hh, 64, 64))                                           # TMEM weight-gradient accumulators
    dirs = np.asarray(tp.scheme.dirs, dtype=np.float64).reshape(n1, tp.n_coords)
    pa, pb = [0, 0, 1, 1, 0, 2], [0, 1, 0, 1, 2, 0]

    def threads():
        for warp in range(8):
            hf, q = warp >> 2, warp & 3
            for lane in range(32):
                ppidx, ug = lane // NUG, lane % NUG
                rowbase = q * 32
                yield dict(warp=warp, lane=lane, hf=hf, q=q, ppidx=ppidx, ug=ug, rowbase=rowbase,
                           p0=rowbase // C + 2 * ppidx, ubase=hf * 32 + ug * UG, R0=rowbase + 2 * C * ppidx)

    def store_rows(img, t, v):                                              # v[pp][c][k]
        own_row = (t["R0"] >> 3) * 1024 + (t["R0"] & 7) * 128
        awr_c = ((t["hf"] * 4 + ((t["ug"] >> 1) if UG == 4 else t["ug"])) ^ (t["R0"] & 7)) << 4
        awr_b = (t["ug"] & 1) * 8 if UG == 4 else 0
        for pp in range(2):
            for c in range(C):
                j = C * pp + c
                dst = own_row + j * 128 + ((awr_c ^ (j << 4)) + awr_b)
                for k in range(UG):
                    t1, t2, t3 = split3(v[pp][c][k])
                    for s, val in enumerate((t1, t2, t3)):
                        img[(s * IMG + dst) // 2 + k] = val

    for tile in range(n_tiles_tc):
        base = tile * T
        slot0 = 0
        for n, nd in enumerate(tp.nets):
            L = len(nd.linears) - 1
            Ws = [np.asarray(p, dtype=np.float64) for p in per_net[n][0::2]]
            m_id = id(nd.module)
            n_out = nd.n_out
            rec = np.full(n_sub * rec_sub_floats, np.nan)
            zimg = np.zeros(3 * IMG // 2)
            aimg = np.zeros(3 * IMG // 2)
            zb_all = {}

            def request_record(h):
                for s in range(n_sub):
                    if base + s * T2 < ws_points:
                        src = zj[base // T2 + s, zj_off[(n, h)]: zj_off[(n, h)] + rec_sub_floats]
                        rec[s * rec_sub_floats:(s + 1) * rec_sub_floats] = src

            def load_record(t, u):
                sub, win = t["p0"] // T2, t["p0"] % T2
                r = sub * rec_sub_floats + u * RS2 + win
                return (np.array([rec[r + c * T2] for c in range(C)]), np.array([rec[r + c * T2 + 1] for c in range(C)]))

            ybar = np.zeros(n_out * C * T)
            for e in range(n_out * C * T):
                row, pt = e // T, e % T
                g = base + pt
                if g < ws_points:
                    ybar[e] = seeds[g // T2, (tp.yrow0[n] + row) * T2 + g % T2]
            request_record(L)
            wlo = Ws[L]                                                     # [n_out][64]

            def wq_of(t):
                gp0 = base + t["p0"]
                live = gp0 < ws_points
                w = np.zeros((2, max(1, WL)))
                if WL and live:
                    for pp in range(2):
                        for d in range(WL):
                            w[pp, d] = wts[gp0 // T2, (n * WL + d) * T2 + t["p0"] % T2 + pp]
                return live, w

            # (1) last Linear
            for t in threads():
                live, wq = wq_of(t)
                zb = np.zeros((2, C, UG))
                for k in range(UG):
                    u = t["ubase"] + k
                    z = load_record(t, u)
                    gbk = 0.0
                    for pp in range(2):
                        ab = np.zeros(C)
                        for o_ in range(n_out):
                            for c in range(C):
                                ab[c] += wlo[o_, u] * ybar[(o_ * C + c) * T + t["p0"] + pp]
                        if live:
                            a, zbk = act_backward(nd.act, z[pp], ab, wq[pp], n1, n2, WL)
                        else:
                            a, zbk = np.zeros(C), np.zeros(C)
                        zb[pp, :, k] = zbk
                        gbk += zbk[0]
                        if live:
                            for o_ in range(n_out):
                                s = sum(ybar[(o_ * C + c) * T + t["p0"] + pp] * a[c] for c in range(C))
                                sgrad[g_wl[n] + o_ * 64 + u] += s
                    sgrad[g_b[(n, L - 1)] + u] += gbk
                zb_all[(t["warp"], t["lane"])] = zb
                if L >= 2:
                    store_rows(zimg, t, zb)
            for o_ in range(n_out):
                sgrad[g_bout[n] + o_] += sum(ybar[(o_ * C) * T + pt] for pt in range(T))

            # (2) hidden layers
            for h in range(L, 1, -1):
                l = h - 1
                slot = slot0 + (l - 1)
                # forward weight images of Linear l: element (row j = out, col k = in) = W_l[j][k], three bf16 terms
                Wsp = split3(Ws[l].astype(np.float32))
                Zs = [decode_image(zimg, s, 128) for s in range(3)]
                D_adj = np.zeros((128, 64))
                for pr in range(6):                                         # D[r][k] = sum_j Z[r][j] W[j][k]
                    D_adj += Zs[pa[pr]] @ np.asarray(Wsp[pb[pr]], dtype=np.float64)
                request_record(h - 1)
                zr_all = {}
                for t in threads():
                    live, wq = wq_of(t)
                    av = np.zeros((2, C, UG))
                    zr = np.zeros((2, C, UG))
                    for k in range(UG):
                        z = load_record(t, t["ubase"] + k)
                        for pp in range(2):
                            zr[pp, :, k] = z[pp]
                            if live:
                                av[pp, :, k] = act_from_record(nd.act, z[pp], wq[pp], n1, n2, WL)
                    zr_all[(t["warp"], t["lane"])] = zr
                    store_rows(aimg, t, av)
                As = [decode_image(aimg, s, 128) for s in range(3)]
                for pr in range(6):                                         # W_bar[j][k] += sum_r Z[r][j] A[r][k]
                    Dw[slot] += Zs[pa[pr]].T @ As[pb[pr]]
                # TMEM row -> staging -> owner
                stage = np.zeros((8, 32 * STAGE_STRIDE))
                for t in threads():
                    row = t["rowbase"] + t["lane"]
                    v = D_adj[row, t["hf"] * 32: t["hf"] * 32 + 32]
                    for s in range(8):
                        stage[t["warp"], t["lane"] * STAGE_STRIDE + 4 * s: t["lane"] * STAGE_STRIDE + 4 * s + 4] = v[4 * s: 4 * s + 4]
                for t in threads():
                    live, wq = wq_of(t)
                    zr = zr_all[(t["warp"], t["lane"])]
                    zb = np.zeros((2, C, UG))
                    gb = np.zeros(UG)
                    for pp in range(2):
                        ab = np.zeros((C, UG))
                        for c in range(C):
                            src = (C * (2 * t["ppidx"] + pp) + c) * STAGE_STRIDE + t["ug"] * UG
                            ab[c] = stage[t["warp"], src: src + UG]
                        for k in range(UG):
                            if live:
                                _, zbk = act_backward(nd.act, zr[pp, :, k], ab[:, k], wq[pp], n1, n2, WL)
                                zb[pp, :, k] = zbk
                            gb[k] += zb[pp, 0, k]
                    for k in range(UG):
                        sgrad[g_b[(n, h - 2)] + t["ubase"] + k] += gb[k]
                    zb_all[(t["warp"], t["lane"])] = zb
                if h > 2:
                    zimg = np.zeros(3 * IMG // 2)
                    for t in threads():
                        store_rows(zimg, t, zb_all[(t["warp"], t["lane"])])

            # (3) Linear 0
            for t in threads():
                zb = zb_all[(t["warp"], t["lane"])]
                gp0 = base + t["p0"]
                for k in range(UG):
                    u = t["ubase"] + k
                    for i in range(nd.widths[0]):
                        ci = nd.in_coord[i]
                        s = sum(zb[pp, 0, k] * coords_all[ci, min(gp0 + pp, N - 1)] for pp in range(2))
                        for f in range(n1):
                            s += (zb[0, 1 + f, k] + zb[1, 1 + f, k]) * dirs[f, ci]
                        sgrad[g_w0[n] + u * nd.widths[0] + i] += s
            slot0 += L - 1

    # ---- final flush ----
    slot = 0
    for n, nd in enumerate(tp.nets):
        L = len(nd.linears) - 1
        m_id = id(nd.module)
        h1, hL, n_in, n_out = nd.widths[1], nd.widths[L], nd.widths[0], nd.n_out
        for e in range(h1 * n_in):
            gpart[theta_off[(m_id, 0)] + e] += sgrad[g_w0[n] + e]
        for hl in range(L):
            for e in range(nd.widths[hl + 1]):
                gpart[theta_off[(m_id, 2 * hl + 1)] + e] += sgrad[g_b[(n, hl)] + e]
        for e in range(n_out * hL):
            o_, k = e // hL, e % hL
            gpart[theta_off[(m_id, 2 * L)] + e] += sgrad[g_wl[n] + o_ * 64 + k]
        for e in range(n_out):
            gpart[theta_off[(m_id, 2 * L + 1)] + e] += sgrad[g_bout[n] + e]
        for l in range(1, L):
            width_j, width_k = nd.widths[l + 1], nd.widths[l]
            tmem = np.zeros((128, 64))                                      # M = 64 accumulator: row j in lane j%16 + 32*(j/16)
            for j in range(64):
                tmem[(j % 16) + 32 * (j // 16)] = Dw[slot, j]
            for warp in range(8):
                hf, q = warp >> 2, warp & 3
                for lane in range(16):
                    j = 16 * q + lane
                    v = tmem[q * 32 + lane, hf * 32: hf * 32 + 32]
                    if j < width_j:
                        for i in range(32):
                            k = hf * 32 + i
                            if k < width_k:
                                gpart[theta_off[(m_id, 2 * l)] + j * width_k + k] += v[i]
            slot += 1

    want = np.concatenate([g.reshape(-1) for g in ref["grads"]])
    err = np.linalg.norm(gpart - want) / np.linalg.norm(want)
    return err, C, (n1, n2, WL), n_tiles_tc


if __name__ == "__main__":
    cases = sys.argv[1:] or ["c2", "c5", "x3"]
    for key in cases:
        for T2 in ((16, 32) if not key.startswith("c5") else (32, 64)):
            try:
                err, C, scheme, nt = emulate(key, 200, T2)
                print(f"{key}: C={C} scheme={scheme} T2={T2} tiles={nt}  rel. gradient error vs the float64 mirror: {err:.3e}")
            except AssertionError as e:
                print(f"{key}: T2={T2}: skipped ({e})")
