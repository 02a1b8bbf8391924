"""TEST INFRASTRUCTURE ONLY (CPU, numpy float64): a line-by-line mirror of the ALGORITHM the CUDA kernels implement.

Where ``oracle/reference_port.py`` restates what the *reference* does (autograd), this file restates what *we* do
(forward Taylor/jet propagation through the FCNN, the traced program, one reverse sweep) so that

* the algebra is validated on the CPU against the golden vectors before any GPU time is spent,
* the kernels' intermediate buffers (z-jets workspace, seeds, per-layer gradients) can be compared 1:1 on the GPU box.

Never imported by the product.
"""
import numpy as np


def act_derivs(act, z0):
    """value and first three derivatives of the activation at z0."""
    if act == 0:  # tanh
        a = np.tanh(z0)
        s1 = 1.0 - a * a
        s2 = -2.0 * a * s1
        s3 = -2.0 * s1 * s1 - 2.0 * a * s2
        return a, s1, s2, s3
    a = np.sin(z0)
    s1 = np.cos(z0)
    return a, s1, -a, -s1


def forward_jets(weights, biases, act, x_in, dirs_in, n2, wl=None):
    """weights[l]: [out,in] (torch layout); x_in: [n_in, N]; dirs_in: [n1, n_in] direction vectors restricted to the
    network inputs.  Channels: 0 value | 1..n1 first order | n1+1..n1+n2 pure second order of the first n2 dirs.
    Returns (z_jets per hidden layer [C, h, N], y [C, n_out, N])."""
    n1 = dirs_in.shape[0]
    if wl is not None:   # combined second-order channel L = sum_d wl[d] * D_d^2  (wl: [n_dirs_weighted, N])
        n2 = 1
    C = 1 + n1 + n2
    N = x_in.shape[1]
    a = np.zeros((C, x_in.shape[0], N))
    a[0] = x_in
    for f in range(n1):
        a[1 + f] = dirs_in[f][:, None]
    z_store = []
    L = len(weights)
    for l in range(L):
        W, b = weights[l], biases[l]
        z = np.einsum("oi,cin->con", W, a)
        z[0] += b[:, None]
        if l == L - 1:
            return z_store, z
        z_store.append(z)
        a = a_from_z(act, z, n1, n2, wl)


def a_from_z(act, z, n1, n2, wl=None):
    a0, s1, s2, _ = act_derivs(act, z[0])
    a = np.empty_like(z)
    a[0] = a0
    for f in range(n1):
        a[1 + f] = s1 * z[1 + f]
    if wl is not None:
        a[1 + n1] = s2 * sum(wl[d] * z[1 + d] ** 2 for d in range(wl.shape[0])) + s1 * z[1 + n1]
        return a
    for s in range(n2):
        a[1 + n1 + s] = s2 * z[1 + s] ** 2 + s1 * z[1 + n1 + s]
    return a


def backward(weights, act, x_in, dirs_in, n2, z_store, ybar, wl=None):
    """ybar: [C, n_out, N] seeds dL/dy.  Returns (grad_W list [out,in], grad_b list)."""
    n1 = dirs_in.shape[0]
    if wl is not None:
        n2 = 1
    C = 1 + n1 + n2
    L = len(weights)
    gW, gb = [None] * L, [None] * L
    zbar = ybar
    for l in range(L - 1, -1, -1):
        if l > 0:
            a_prev = a_from_z(act, z_store[l - 1], n1, n2, wl)
        else:
            a_prev = np.zeros((C, x_in.shape[0], x_in.shape[1]))
            a_prev[0] = x_in
            for f in range(n1):
                a_prev[1 + f] = dirs_in[f][:, None]
        gW[l] = np.einsum("con,cin->oi", zbar, a_prev)
        gb[l] = zbar[0].sum(axis=1)
        if l == 0:
            break
        abar = np.einsum("oi,con->cin", weights[l], zbar)
        z = z_store[l - 1]
        _, s1, s2, s3 = act_derivs(act, z[0])
        zb = np.empty_like(z)
        zb[0] = s1 * abar[0]
        for f in range(n1):
            zb[1 + f] = s1 * abar[1 + f]
            zb[0] += s2 * z[1 + f] * abar[1 + f]
        if wl is not None:
            aL = abar[1 + n1]
            zb[1 + n1] = s1 * aL
            q = 0.0
            for d in range(wl.shape[0]):
                zb[1 + d] += 2.0 * s2 * wl[d] * z[1 + d] * aL
                q = q + wl[d] * z[1 + d] ** 2
            zb[0] += (s3 * q + s2 * z[1 + n1]) * aL
        else:
            for s in range(n2):
                zb[1 + n1 + s] = s1 * abar[1 + n1 + s]
                zb[1 + s] += 2.0 * s2 * z[1 + s] * abar[1 + n1 + s]
                zb[0] += (s3 * z[1 + s] ** 2 + s2 * z[1 + n1 + s]) * abar[1 + n1 + s]
        zbar = zb
    return gW, gb


def run_traced(tp, params_per_net, coords, n_global=None, want_grad=True, rbar=None, ubar=None):
    """Evaluate a TracedProblem end to end in float64.

    ``params_per_net``: list (per network INSTANCE of ``tp.nets``; instances of one module get the same arrays) of
    [W0,b0,W1,b1,...] numpy arrays (torch layout).
    Returns dict(u, residual, loss, grads (flat list per distinct MODULE, instances summed), y, seeds, z_store)."""
    from neurodiffeq_b200 import symbolic as S
    coords = tp.extend_coords(np.asarray(coords, dtype=np.float64))   # + constant coordinates (boundary instances)
    N = coords.shape[1]
    dirs = np.asarray(tp.scheme.dirs, dtype=np.float64).reshape(tp.scheme.n1, tp.n_coords)
    n1, n2 = tp.scheme.n1, tp.scheme.n2
    C = tp.n_channels
    theta = {}      # trainable scalars that enter the residual program directly: Resnet shortcut matrices
    for k, nd in enumerate(tp.nets):
        if getattr(nd, "skip", None) is not None:
            w_skip = np.asarray(params_per_net[k][2 * len(nd.linears)], dtype=np.float64)
            for o in range(w_skip.shape[0]):
                for i in range(w_skip.shape[1]):
                    theta[("skip", id(nd.module), o, i)] = w_skip[o, i]
    wl_all = None
    if getattr(tp, "wl", 0):
        wl_all = S.evaluate_program(tp.prog_w, coords, np.zeros((1, N)), n_w=len(tp.nets) * tp.wl, theta=theta)
    y_rows = np.zeros((tp.n_yrows, N))
    stores = []
    for k, nd in enumerate(tp.nets):
        wl = wl_all[k * tp.wl:(k + 1) * tp.wl] if wl_all is not None else None
        body = params_per_net[k][:2 * len(nd.linears)]
        Ws = [np.asarray(p, dtype=np.float64) for p in body[0::2]]
        bs = [np.asarray(p, dtype=np.float64) for p in body[1::2]]
        x_in = coords[list(nd.in_coord)]
        d_in = dirs[:, list(nd.in_coord)]
        z_store, y = forward_jets(Ws, bs, nd.act, x_in, d_in, n2, wl)
        stores.append((Ws, x_in, d_in, z_store, wl))
        for o in range(nd.n_out):
            for c in range(C):
                y_rows[tp.yrow0[k] + o * C + c] = y[c, o]
    u, r, _ = S.evaluate_program(tp.prog_eval, coords, y_rows, n_u=tp.n_funcs, n_r=tp.n_eq, theta=theta)
    out = dict(u=u, residual=r, loss=float((r ** 2).mean()) if r.size else 0.0, y=y_rows)
    if want_grad:
        n_glob = N if n_global is None else n_global
        scale = 2.0 / (n_glob * tp.n_eq)
        if rbar is None:   # L = mean(r^2) over the global batch
            _, r2, seeds = S.evaluate_program(tp.prog_train, coords, y_rows, params=[scale], n_r=tp.n_eq,
                                              n_seed=tp.n_yrows, theta=theta)
        elif ubar is None:  # externally supplied dL/dr [n_eq, N] (custom loss functions)
            _, r2, seeds = S.evaluate_program(tp.prog_train_ext, coords, y_rows, rbar=np.asarray(rbar, dtype=np.float64),
                                              params=[scale], n_r=tp.n_eq, n_seed=tp.n_yrows, theta=theta)
        else:               # ... and dL/du [n_funcs, N] for losses that also depend on the functions
            ext = np.concatenate([np.asarray(rbar, dtype=np.float64), np.asarray(ubar, dtype=np.float64)], axis=0)
            _, r2, seeds = S.evaluate_program(tp.prog_train_ext_u, coords, y_rows, rbar=ext, params=[scale], n_r=tp.n_eq,
                                              n_seed=tp.n_yrows, theta=theta)
        assert np.allclose(r2, r)
        by_module = {}   # instances that share a module (network evaluated at a boundary too) add up, like autograd
        for k, nd in enumerate(tp.nets):
            Ws, x_in, d_in, z_store, wl = stores[k]
            ybar = np.zeros((C, nd.n_out, N))
            for o in range(nd.n_out):
                for c in range(C):
                    ybar[c, o] = seeds[tp.yrow0[k] + o * C + c]
            gW, gb = backward(Ws, nd.act, x_in, d_in, n2, z_store, ybar, wl)
            mine = []
            for w, b in zip(gW, gb):
                mine += [w, b]
            if getattr(nd, "skip", None) is not None:
                # shortcut matrix of a Resnet: the raw output is (network jet + shortcut jet), so both share the seeds;
                # d(value)/dW_s[o][i] = x_i, d(first-order channel f)/dW_s[o][i] = dir_f[i], second-order channels: 0
                g_skip = np.zeros((nd.n_out, len(nd.in_coord)))
                for o in range(nd.n_out):
                    for i in range(len(nd.in_coord)):
                        g_skip[o, i] = (ybar[0, o] * x_in[i]).sum() + sum(ybar[1 + f, o].sum() * d_in[f, i] for f in range(n1))
                mine.append(g_skip)
            acc = by_module.setdefault(id(nd.module), mine)
            if acc is not mine:
                for a, m in zip(acc, mine):
                    a += m
        grads = [g_ for gs in by_module.values() for g_ in gs]
        out.update(grads=grads, seeds=seeds, z_store=[s[3] for s in stores])
    return out
