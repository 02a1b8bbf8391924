"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; never imported by ``neurodiffeq_b200``.

A from-scratch restatement, in plain PyTorch autograd on the CPU, of the one hot path of the reference
(NeuroDiffGym/neurodiffeq @ 9f6d6e3): collocation batch -> FCNN -> condition re-parameterisation -> PDE residual via
repeated ``torch.autograd.grad(create_graph=True)`` -> mean-squared loss -> ``loss.backward()``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import this file, and only as the checker / the CPU arm.  The arithmetic lives in PyTorch (a third-party dependency of
the reference, ``requirements.txt:5``; version here: torch 2.11.0); parity is PINNED by the golden vectors in
``tests/golden/*.npz`` which were produced by the unmodified reference itself (``tests/golden/generate.py``) and
which ``tests/test_oracle.py`` compares this file against (residual, loss and every parameter gradient).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import types

import torch
import torch.nn as nn


# ----------------------------------------------------------------------------------------------------------------------
# diff                                                                     neurodiffeq/neurodiffeq.py:6-82
# ----------------------------------------------------------------------------------------------------------------------
def _one_derivative(u, t):
    """d u / d t for per-sample-independent rows: grad with ones as cotangent; unused input -> zeros (:22-24)."""
    g, = torch.autograd.grad(u, t, grad_outputs=torch.ones_like(u), create_graph=True, allow_unused=True)
    if g is None:
        return None
    return g.requires_grad_()


def diff(u, t, order=1, shape_check=True):
    if shape_check:  # :52-59
        if u.dim() != 2 or t.dim() != 2 or u.shape[1] != 1 or t.shape[1] != 1:
            raise ValueError(f"Input shapes must both be (n_samples, 1); got {u.shape} and {t.shape}")
        if u.shape != t.shape:
            raise ValueError(f"Input shapes must be the same; got {u.shape} != {t.shape}")
    cur = u
    for _ in range(order):  # :21-33
        cur = _one_derivative(cur, t)
        if cur is None:
            return torch.zeros_like(t, requires_grad=True)
    return cur


# ----------------------------------------------------------------------------------------------------------------------
# operators                                                                neurodiffeq/operators.py:15-207
# ----------------------------------------------------------------------------------------------------------------------
def grad(u, *xs):  # :15-33 (one autograd.grad call for all coordinates)
    gs = torch.autograd.grad(u, xs, grad_outputs=torch.ones_like(u), create_graph=True, allow_unused=True)
    return [torch.zeros_like(x, requires_grad=True) if g is None else g.requires_grad_(True) for x, g in zip(xs, gs)]


def div(*us_xs):  # :36-49
    n = len(us_xs)
    if n == 0 or n % 2:
        raise RuntimeError("Number of us and xs must be equal and positive")
    return sum(diff(u, x) for u, x in zip(us_xs[:n // 2], us_xs[n // 2:]))


def curl(u_x, u_y, u_z, x, y, z):  # :52-74
    uxy, uxz = grad(u_x, y, z)
    uyx, uyz = grad(u_y, x, z)
    uzx, uzy = grad(u_z, x, y)
    return uzy - uyz, uxz - uzx, uyx - uxy


def laplacian(u, *xs):  # :77-89
    return sum(diff(g, x) for g, x in zip(grad(u, *xs), xs))


def spherical_laplacian(u, r, theta, phi):  # :189-207
    u_r, u_th, u_ph = grad(u, r, theta, phi)
    s = torch.sin(theta)
    rr = r ** 2
    return (diff(rr * u_r, r) + diff(s * u_th, theta) / s + diff(u_ph, phi) / s ** 2) / rr


# ----------------------------------------------------------------------------------------------------------------------
# networks                                                                 neurodiffeq/networks.py:6-70, 142-152
# ----------------------------------------------------------------------------------------------------------------------
class SinActv(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class FCNN(nn.Module):
    """Linear/activation stack ending in a Linear; parameters live at NN.{0,2,4,...}.{weight,bias} like the
    reference so that state_dict order (and hence golden parameter order) is identical."""

    def __init__(self, n_input_units=1, n_output_units=1, actv=nn.Tanh, hidden_units=(32, 32)):
        super().__init__()
        widths = (n_input_units,) + tuple(hidden_units)
        mods = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(a, b), actv()]
        mods.append(nn.Linear(widths[-1], n_output_units))
        self.NN = nn.Sequential(*mods)

    def forward(self, x):
        return self.NN(x)


class Resnet(nn.Module):
    """networks.py:73-106: ``skip_connection(t) + residual(t)`` with a bias-free Linear shortcut"""

    def __init__(self, n_input_units=1, n_output_units=1, actv=nn.Tanh, hidden_units=(32, 32)):
        super().__init__()
        self.residual = FCNN(n_input_units=n_input_units, n_output_units=n_output_units, actv=actv, hidden_units=hidden_units)
        self.skip_connection = nn.Linear(n_input_units, n_output_units, bias=False)

    def forward(self, x):
        return self.skip_connection(x) + self.residual(x)


# ----------------------------------------------------------------------------------------------------------------------
# conditions                                                               neurodiffeq/conditions.py
# ----------------------------------------------------------------------------------------------------------------------
class _Condition:
    def __init__(self):
        self.ith_unit = None

    def set_impose_on(self, i):  # :59-75
        self.ith_unit = i

    def enforce(self, net, *coords):  # :41-57
        out = net(torch.cat(coords, dim=1))
        if self.ith_unit is not None:
            out = out[:, self.ith_unit].view(-1, 1)
        return self.parameterize(out, *coords)


class NoCondition(_Condition):  # :205-222
    def parameterize(self, out, *coords):
        return out


class EnsembleCondition(_Condition):  # :157-202
    def __init__(self, *sub_conditions, force=False):
        super().__init__()
        self.conditions = sub_conditions

    def parameterize(self, out, *coords):  # column i of the network output goes through sub-condition i
        if out.shape[1] != len(self.conditions):
            raise ValueError(f"number of output units ({out.shape[1]}) differs from number of conditions "
                             f"({len(self.conditions)})")
        return torch.cat([c.parameterize(out[:, i].view(-1, 1), *coords) for i, c in enumerate(self.conditions)], dim=1)


class IVP(_Condition):  # :225-267
    def __init__(self, t_0, u_0=None, u_0_prime=None):
        super().__init__()
        self.t_0, self.u_0, self.u_0_prime = t_0, u_0, u_0_prime

    def parameterize(self, out, t):
        decay = 1 - torch.exp(-t + self.t_0)
        if self.u_0_prime is None:
            return self.u_0 + decay * out
        return self.u_0 + (t - self.t_0) * self.u_0_prime + decay ** 2 * out


class BundleIVP(_Condition):  # :270-345 with the lookup of :109-135
    def __init__(self, t_0=None, u_0=None, u_0_prime=None, bundle_param_lookup=None):
        super().__init__()
        self.t_0, self.u_0, self.u_0_prime = t_0, u_0, u_0_prime
        self.bundle_param_lookup = bundle_param_lookup or {}

    def _lookup(self, name, theta):
        if name in self.bundle_param_lookup:
            return theta[self.bundle_param_lookup[name]]
        return getattr(self, name)

    def parameterize(self, out, t, *theta):
        t_0, u_0, u_0p = (self._lookup(k, theta) for k in ("t_0", "u_0", "u_0_prime"))
        decay = 1 - torch.exp(-t + t_0)
        if u_0p is None:
            return u_0 + decay * out
        return u_0 + (t - t_0) * u_0p + decay ** 2 * out


class DirichletBVP2D(_Condition):  # :438-509
    def __init__(self, x_min, x_min_val, x_max, x_max_val, y_min, y_min_val, y_max, y_max_val):
        super().__init__()
        self.x0, self.f0, self.x1, self.f1 = x_min, x_min_val, x_max, x_max_val
        self.y0, self.g0, self.y1, self.g1 = y_min, y_min_val, y_max, y_max_val

    def parameterize(self, out, x, y):
        xt = (x - self.x0) / (self.x1 - self.x0)
        yt = (y - self.y0) / (self.y1 - self.y0)
        x0 = torch.full_like(xt, self.x0)
        x1 = torch.full_like(xt, self.x1)
        lin = lambda g: (1 - xt) * g(x0) + xt * g(x1)  # noqa: E731  corner correction of :505-507
        a = (1 - xt) * self.f0(y) + xt * self.f1(y) + (1 - yt) * (self.g0(x) - lin(self.g0)) \
            + yt * (self.g1(x) - lin(self.g1))
        return a + xt * (1 - xt) * yt * (1 - yt) * out


def _ann(cond, net, *cols):  # the local ``ANN`` helper of conditions.py:577-581 / :816-820
    out = net(torch.cat(cols, dim=1))
    if cond.ith_unit is not None:
        out = out[:, cond.ith_unit].view(-1, 1)
    return out


def _which_ends(cond):  # branch selection by truthiness (IBVP1D :583-600) / by `is not None` (DoubleEndedBVP1D :822-840)
    return ("d" if cond.lo_val_given else "n") + ("d" if cond.hi_val_given else "n")


class IBVP1D(_Condition):  # :512-712 (enforce :559-600, parameterize :603-712)
    def __init__(self, x_min, x_max, t_min, t_min_val, x_min_val=None, x_min_prime=None, x_max_val=None,
                 x_max_prime=None):
        super().__init__()
        given = [c is not None for c in (x_min_val, x_min_prime, x_max_val, x_max_prime)]
        if sum(given) != 2 or (x_min_val and x_min_prime) or (x_max_val and x_max_prime):   # :546-548
            raise NotImplementedError("Sorry, this boundary condition is not implemented.")
        self.x_min, self.x_max, self.t_min, self.t_min_val = x_min, x_max, t_min, t_min_val
        self.x_min_val, self.x_min_prime, self.x_max_val, self.x_max_prime = x_min_val, x_min_prime, x_max_val, x_max_prime
        self.lo_val_given, self.hi_val_given = bool(x_min_val), bool(x_max_val)

    def enforce(self, net, x, t):
        u = _ann(self, net, x, t)
        extra = []
        kind = _which_ends(self)
        if kind[0] == "n":   # network at the left boundary, on a fresh leaf so that d/dx0 can be taken (:590, :594)
            x0 = self.x_min * torch.ones_like(x, requires_grad=True)
            extra += [_ann(self, net, x0, t), x0]
        if kind[1] == "n":   # ... and at the right boundary (:586, :595)
            x1 = self.x_max * torch.ones_like(x, requires_grad=True)
            extra += [_ann(self, net, x1, t), x1]
        return self.parameterize(u, x, t, *extra)

    def parameterize(self, out, x, t, *extra):
        t0 = self.t_min * torch.ones_like(t, requires_grad=True)
        s = (x - self.x_min) / (self.x_max - self.x_min)          # x tilde
        tau = t - self.t_min                                       # t tilde
        width = self.x_max - self.x_min
        fade = 1 - torch.exp(-tau)
        kind = _which_ends(self)
        if kind == "dd":    # :661-666
            a = self.t_min_val(x) + s * (self.x_max_val(t) - self.x_max_val(t0)) \
                + (1 - s) * (self.x_min_val(t) - self.x_min_val(t0))
            return a + s * (1 - s) * fade * out
        if kind == "dn":    # :670-676
            n1, x1 = extra
            a = (self.x_min_val(t) - self.x_min_val(t0)) + self.t_min_val(x) \
                + s * width * (self.x_max_prime(t) - self.x_max_prime(t0))
            return a + s * fade * (out - width * diff(n1, x1) - n1)
        if kind == "nd":    # :680-686
            n0, x0 = extra
            a = (self.x_max_val(t) - self.x_max_val(t0)) + self.t_min_val(x) \
                + (s - 1) * width * (self.x_min_prime(t) - self.x_min_prime(t0))
            return a + (1 - s) * fade * (out + width * diff(n0, x0) - n0)
        n0, x0, n1, x1 = extra   # :689-701
        a = self.t_min_val(x) - 0.5 * (1 - s) ** 2 * width * (self.x_min_prime(t) - self.x_min_prime(t0)) \
            + 0.5 * s ** 2 * width * (self.x_max_prime(t) - self.x_max_prime(t0))
        d0, d1 = diff(n0, x0), diff(n1, x1)
        return a + fade * (out - s * width * d0 + 0.5 * s ** 2 * width * (d0 - d1))


class DoubleEndedBVP1D(_Condition):  # :715-883 (enforce :797-840, formulas of the CODE :857-883)
    def __init__(self, x_min, x_max, x_min_val=None, x_min_prime=None, x_max_val=None, x_max_prime=None):
        super().__init__()
        given = [c is not None for c in (x_min_val, x_min_prime, x_max_val, x_max_prime)]
        if sum(given) != 2 or (x_min_val and x_min_prime) or (x_max_val and x_max_prime):   # :751-753
            raise NotImplementedError("Sorry, this boundary condition is not implemented.")
        self.x_min, self.x_max = x_min, x_max
        self.x_min_val, self.x_min_prime, self.x_max_val, self.x_max_prime = x_min_val, x_min_prime, x_max_val, x_max_prime
        self.lo_val_given, self.hi_val_given = x_min_val is not None, x_max_val is not None

    def enforce(self, net, x):
        u = _ann(self, net, x)
        extra = []
        kind = _which_ends(self)
        if kind[0] == "n":
            x0 = self.x_min * torch.ones_like(x, requires_grad=True)
            extra += [_ann(self, net, x0), x0]
        if kind[1] == "n":
            x1 = self.x_max * torch.ones_like(x, requires_grad=True)
            extra += [_ann(self, net, x1), x1]
        return self.parameterize(u, x, *extra)

    def parameterize(self, out, x, *extra):
        s = (x - self.x_min) / (self.x_max - self.x_min)
        width = self.x_max - self.x_min
        kind = _which_ends(self)
        if kind == "dd":    # :857-859
            return self.x_min_val * (1 - s) + self.x_max_val * s + s * (1 - s) * out
        if kind == "dn":    # :863-865
            n1, x1 = extra
            a = (1 - s) * self.x_min_val + 0.5 * s ** 2 * self.x_max_prime * width
            return a + s * (out - n1 + self.x_min_val - diff(n1, x1) * width)
        if kind == "nd":    # :871-873
            n0, x0 = extra
            a = s * self.x_max_val - 0.5 * (1 - s) ** 2 * self.x_min_prime * width
            return a + (1 - s) * (out - n0 + self.x_max_val + diff(n0, x0) * width)
        n0, x0, n1, x1 = extra   # :878-882
        a = -0.5 * (1 - s) ** 2 * width * self.x_min_prime + 0.5 * s ** 2 * width * self.x_max_prime
        return a + 0.5 * s ** 2 * (out - n1 - 0.5 * diff(n1, x1) * width) \
            + 0.5 * (1 - s) ** 2 * (out - n0 + 0.5 * diff(n0, x0) * width)


class DirichletBVPSpherical(_Condition):  # :887-956
    def __init__(self, r_0, f, r_1=None, g=None):
        super().__init__()
        if (r_1 is None) != (g is None):
            raise ValueError("r_1 and g must be both/neither None")
        self.r_0, self.f, self.r_1, self.g = r_0, f, r_1, g

    def parameterize(self, out, r, theta, phi):
        if self.r_1 is None:
            return (1 - torch.exp(-torch.abs(r - self.r_0))) * out + self.f(theta, phi)
        rt = (r - self.r_0) / (self.r_1 - self.r_0)
        return self.f(theta, phi) * (1 - rt) + self.g(theta, phi) * rt + (1. - torch.exp((1 - rt) * rt)) * out


NAMESPACE = types.SimpleNamespace(
    diff=diff, grad=grad, div=div, curl=curl, laplacian=laplacian, spherical_laplacian=spherical_laplacian,
    FCNN=FCNN, Resnet=Resnet, SinActv=SinActv, NoCondition=NoCondition, EnsembleCondition=EnsembleCondition, IVP=IVP, BundleIVP=BundleIVP,
    DirichletBVP2D=DirichletBVP2D, IBVP1D=IBVP1D, DoubleEndedBVP1D=DoubleEndedBVP1D,
    DirichletBVPSpherical=DirichletBVPSpherical)


# ----------------------------------------------------------------------------------------------------------------------
# the closure                                                              neurodiffeq/solvers.py:369-395
# ----------------------------------------------------------------------------------------------------------------------
def distinct_modules(nets):
    seen, out = set(), []
    for n in nets:
        if id(n) not in seen:
            seen.add(id(n))
            out.append(n)
    return out


def closure(nets, conditions, diff_eqs, coords, backward=True):
    """coords: list of (N,1) tensors with requires_grad.  Returns (funcs, residual (N,n_eq), loss).
    Gradients accumulate into ``p.grad`` like the reference (zero_grad is the caller's business, :360-362)."""
    funcs = [c.enforce(n, *coords) for n, c in zip(nets, conditions)]  # :373-375
    residual = torch.cat(diff_eqs(*funcs, *coords), dim=1)  # :380-381
    loss = (residual ** 2).mean()  # :218
    if backward:
        loss.backward()  # :393
    return funcs, residual, loss


def evaluate(nets, conditions, diff_eqs, coords_soa, dtype=torch.float64, backward=True):
    """numpy-in / numpy-out convenience around :func:`closure` for tests and the CPU baseline.
    ``coords_soa``: array [d0, N].  Returns dict(u [n_funcs,N], residual [n_eq,N], loss, grads [per param])."""
    mods = distinct_modules(nets)
    for m in mods:
        m.to(dtype)
        for p in m.parameters():
            p.grad = None
    coords = [torch.as_tensor(c, dtype=dtype).reshape(-1, 1).requires_grad_(True) for c in coords_soa]
    funcs, residual, loss = closure(nets, conditions, diff_eqs, coords, backward=backward)
    out = dict(u=torch.cat([f.detach().t() for f in funcs]).numpy(), residual=residual.detach().numpy().T.copy(),   # ensemble: k rows
               loss=float(loss.detach()))
    if backward:
        out["grads"] = [p.grad.detach().numpy().copy() for m in mods for p in m.parameters()]
    return out


def load_params(nets, arrays, dtype=torch.float64):
    """Copy a list of numpy arrays (state_dict order over the distinct nets) into the modules."""
    it = iter(arrays)
    for m in distinct_modules(nets):
        m.to(dtype)
        for p in m.parameters():
            p.data = torch.as_tensor(next(it), dtype=dtype).reshape(p.shape).clone()
