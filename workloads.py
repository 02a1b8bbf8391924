"""BASELINE.json workloads C1..C5, written ONCE against an abstract namespace ``nd``.

The same builder runs against three namespaces that expose the reference's names
(``FCNN, SinActv, IVP, BundleIVP, DirichletBVP2D, IBVP1D, DirichletBVPSpherical, diff, spherical_laplacian``):

* the unmodified reference (``tools/ref_shim.py``; only in the build container, to make ``tests/golden``),
* the CPU oracle (``oracle/``; tests + ``bench.py`` baseline legs only),
* the product (``neurodiffeq_b200``; traces the same callables symbolically and runs the CUDA kernels).

so that parity tests read like the reference's own usage (README.md:74-130 of the reference).
Coordinates are synthetic (``numpy.random.RandomState(seed)``, fp32-representable values) so that every
implementation sees bit-identical inputs; sampling by the generator classes is tested separately.
"""
import math
from collections import namedtuple

import numpy as np
import torch

Workload = namedtuple(
    "Workload",
    "name solver coord_names coord_ranges nets_spec make_nets make_conditions diff_eqs n_eq default_n flops_fwdjet "
    "eq_param_index"
)

NU_BURGERS = 0.01 / math.pi


def _fcnn_flops(widths, n_channels):
    """F_fwdjet = 2*d0*h1 + C*2*(sum_{l>=2} h_{l-1} h_l + h_L*dout)."""
    d0, h = widths[0], widths[1:]
    rest = sum(a * b for a, b in zip(h[:-1], h[1:]))
    return 2 * d0 * h[0] + n_channels * 2 * rest


# ----------------------------------------------------------------------------------------------------------------------
# C1  Solver1D Lotka-Volterra, 2 x FCNN(1-32-32-1, SinActv)            (reference README.md:85-93)
# ----------------------------------------------------------------------------------------------------------------------
def _c1(nd):
    def make_nets():
        return [nd.FCNN(n_input_units=1, n_output_units=1, hidden_units=(32, 32), actv=nd.SinActv) for _ in range(2)]

    def make_conditions():
        return [nd.IVP(t_0=0.0, u_0=1.5), nd.IVP(t_0=0.0, u_0=1.0)]

    def diff_eqs(u, v, t):
        return [nd.diff(u, t) - (u - u * v), nd.diff(v, t) - (u * v - v)]

    return Workload("c1_lotka_volterra", "Solver1D", ("t",), ((0.1, 12.0),),
                    [((1, 32, 32, 1), "sin")] * 2, make_nets, make_conditions, diff_eqs, 2, 1024,
                    2 * _fcnn_flops((1, 32, 32, 1), 2), None)


# ----------------------------------------------------------------------------------------------------------------------
# C2  Solver2D Laplace, DirichletBVP2D, FCNN(2-64-64-64-1, tanh)        (reference README.md:113-128)
# ----------------------------------------------------------------------------------------------------------------------
def _c2(nd):
    def make_nets():
        return [nd.FCNN(n_input_units=2, n_output_units=1, hidden_units=(64, 64, 64))]

    def make_conditions():
        return [nd.DirichletBVP2D(
            x_min=0, x_min_val=lambda y: torch.sin(np.pi * y),
            x_max=1, x_max_val=lambda y: 0,
            y_min=0, y_min_val=lambda x: 0,
            y_max=1, y_max_val=lambda x: 0,
        )]

    def diff_eqs(u, x, y):
        return [nd.diff(u, x, order=2) + nd.diff(u, y, order=2)]

    return Workload("c2_laplace2d", "Solver2D", ("x", "y"), ((0.0, 1.0), (0.0, 1.0)),
                    [((2, 64, 64, 64, 1), "tanh")], make_nets, make_conditions, diff_eqs, 1, 16384,
                    _fcnn_flops((2, 64, 64, 64, 1), 5), None)


# ----------------------------------------------------------------------------------------------------------------------
# C3  Solver2D Burgers, IBVP1D (Dirichlet-Dirichlet), FCNN(2-128-128-128-1, tanh)     
# ----------------------------------------------------------------------------------------------------------------------
def _c3(nd):
    def make_nets():
        return [nd.FCNN(n_input_units=2, n_output_units=1, hidden_units=(128, 128, 128))]

    def make_conditions():
        return [nd.IBVP1D(
            x_min=-1, x_max=1, t_min=0,
            t_min_val=lambda x: -torch.sin(np.pi * x),
            x_min_val=lambda t: 0,
            x_max_val=lambda t: 0,
        )]

    def diff_eqs(u, x, t):
        return [nd.diff(u, t) + u * nd.diff(u, x) - NU_BURGERS * nd.diff(u, x, order=2)]

    return Workload("c3_burgers", "Solver2D", ("x", "t"), ((-1.0, 1.0), (0.0, 1.0)),
                    [((2, 128, 128, 128, 1), "tanh")], make_nets, make_conditions, diff_eqs, 1, 65536,
                    _fcnn_flops((2, 128, 128, 128, 1), 4), None)


# ----------------------------------------------------------------------------------------------------------------------
# C4  SolverSpherical Poisson with spherical_laplacian, FCNN(3-64-64-64-1)  (reference tests/test_pde_spherical.py:103)
# ----------------------------------------------------------------------------------------------------------------------
def _c4(nd):
    r0, r1 = 0.1, 3.0
    k_q = 1.0 / (4 * math.pi)
    v0 = k_q / r0 * math.erf(r0 / math.sqrt(2))
    v1 = k_q / r1 * math.erf(r1 / math.sqrt(2))
    norm = (2 * math.pi) ** 1.5

    def make_nets():
        return [nd.FCNN(n_input_units=3, n_output_units=1, hidden_units=(64, 64, 64))]

    def make_conditions():
        return [nd.DirichletBVPSpherical(r_0=r0, f=lambda th, ph: v0, r_1=r1, g=lambda th, ph: v1)]

    def diff_eqs(u, r, th, ph):
        return [nd.spherical_laplacian(u, r, th, ph) + torch.exp(-r ** 2 / 2) / norm]

    return Workload("c4_spherical_poisson", "SolverSpherical", ("r", "theta", "phi"),
                    ((r0, r1), (0.07, math.pi - 0.07), (0.0, 2 * math.pi)),
                    [((3, 64, 64, 64, 1), "tanh")], make_nets, make_conditions, diff_eqs, 1, 32768,
                    _fcnn_flops((3, 64, 64, 64, 1), 7), None)


# ----------------------------------------------------------------------------------------------------------------------
# C5  BundleSolver1D damped oscillator, 4 bundle params, ONE shared FCNN(5-64-64-2) with ith_unit  
# ----------------------------------------------------------------------------------------------------------------------
def _c5(nd):
    def make_nets():
        net = nd.FCNN(n_input_units=5, n_output_units=2, hidden_units=(64, 64))
        return [net, net]

    def make_conditions():
        import warnings
        conds = [nd.BundleIVP(t_0=0.0, bundle_param_lookup={"u_0": 2}),
                 nd.BundleIVP(t_0=0.0, bundle_param_lookup={"u_0": 3})]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, c in enumerate(conds):
                c.set_impose_on(i)
        return conds

    def diff_eqs(u, v, t, zeta, omega):
        return [nd.diff(u, t) - v, nd.diff(v, t) + 2 * zeta * omega * v + omega ** 2 * u]

    return Workload("c5_bundle_oscillator", "BundleSolver1D", ("t", "zeta", "omega", "u0", "v0"),
                    ((0.0, 2 * math.pi), (0.05, 0.5), (0.5, 2.0), (-1.0, 1.0), (-1.0, 1.0)),
                    [((5, 64, 64, 2), "tanh")], make_nets, make_conditions, diff_eqs, 2, 131072,
                    _fcnn_flops((5, 64, 64, 2), 2), (0, 1))


# ----------------------------------------------------------------------------------------------------------------------
# Extension workloads: conditions with Neumann data, which evaluate the network AT a boundary abscissa
# (reference conditions.py:585-596, 823-834).  Not BASELINE configs -- parity cases for the widened condition family.
# ----------------------------------------------------------------------------------------------------------------------
def _heat(nd, name, neumann_side):
    def make_nets():
        return [nd.FCNN(n_input_units=2, n_output_units=1, hidden_units=(64, 64))]

    def make_conditions():
        kw = dict(x_min=0.0, x_max=1.0, t_min=0.0, t_min_val=lambda x: torch.sin(0.5 * np.pi * x))
        if neumann_side == "right":   # Dirichlet at x0, Neumann at x1 (conditions.py:670-676)
            kw.update(x_min_val=lambda t: 0.2 * torch.sin(t), x_max_prime=lambda t: 0.1 * t)
        elif neumann_side == "left":  # Neumann at x0, Dirichlet at x1 (conditions.py:680-686)
            kw.update(x_min_prime=lambda t: 0.1 * t, x_max_val=lambda t: 1.0 + 0.2 * torch.sin(t))
        else:                         # Neumann at both ends (conditions.py:689-701)
            kw.update(x_min_prime=lambda t: 0.5 * np.pi + 0.1 * t, x_max_prime=lambda t: 0.3 * torch.sin(t))
        return [nd.IBVP1D(**kw)]

    def diff_eqs(u, x, t):
        return [nd.diff(u, t) - 0.3 * nd.diff(u, x, order=2)]

    return Workload(name, "Solver2D", ("x", "t"), ((0.0, 1.0), (0.0, 1.0)), [((2, 64, 64, 1), "tanh")], make_nets,
                    make_conditions, diff_eqs, 1, 16384, 2 * _fcnn_flops((2, 64, 64, 1), 9), None)


def _bvp(nd, name, kw):
    def make_nets():
        return [nd.FCNN(n_input_units=1, n_output_units=1, hidden_units=(32, 32))]

    def make_conditions():
        return [nd.DoubleEndedBVP1D(0.0, 1.0, **kw)]

    def diff_eqs(u, x):
        return [nd.diff(u, x, order=2) + u - x]

    n_inst = 1 + sum(k.endswith("prime") for k in kw)
    return Workload(name, "Solver1D", ("x",), ((0.0, 1.0),), [((1, 32, 32, 1), "tanh")], make_nets, make_conditions,
                    diff_eqs, 1, 4096, n_inst * _fcnn_flops((1, 32, 32, 1), 3), None)


def _ensemble(nd):
    """One 2-output network, EnsembleCondition of two IVPs (reference conditions.py:157-202): u' = v, v' = -u."""
    def make_nets():
        return [nd.FCNN(n_input_units=1, n_output_units=2, hidden_units=(32, 32), actv=nd.SinActv)]

    def make_conditions():
        return [nd.EnsembleCondition(nd.IVP(t_0=0.0, u_0=0.0), nd.IVP(t_0=0.0, u_0=1.0))]

    def diff_eqs(uv, t):
        u, v = uv[:, 0:1], uv[:, 1:2]
        return [nd.diff(u, t) - v, nd.diff(v, t) + u]

    return Workload("x7_ensemble_oscillator", "Solver1D", ("t",), ((0.0, 6.0),), [((1, 32, 32, 2), "sin")], make_nets,
                    make_conditions, diff_eqs, 2, 4096, _fcnn_flops((1, 32, 32, 2), 2), None)


def _resnet(nd):
    """Resnet (FCNN + bias-free shortcut, reference networks.py:73-106) on Burgers-like dynamics in (x, t)."""
    def make_nets():
        return [nd.Resnet(n_input_units=2, n_output_units=1, hidden_units=(32, 32))]

    def make_conditions():
        return [nd.IBVP1D(x_min=-1, x_max=1, t_min=0, t_min_val=lambda x: -torch.sin(np.pi * x),
                          x_min_val=lambda t: 0, x_max_val=lambda t: 0)]

    def diff_eqs(u, x, t):
        return [nd.diff(u, t) + u * nd.diff(u, x) - 0.05 * nd.diff(u, x, order=2)]

    return Workload("x9_resnet_burgers", "Solver2D", ("x", "t"), ((-1.0, 1.0), (0.0, 1.0)), [((2, 32, 32, 1), "tanh")],
                    make_nets, make_conditions, diff_eqs, 1, 4096, _fcnn_flops((2, 32, 32, 1), 4), None)


def _third_order(nd):
    """u(3) + u(1) = 0, u(0) = 1, u(1)(0) = 0: order-3 derivative of a network output -- beyond the order-2 jets of the fused
    kernels (the reference nests diff to order 3 and more: tests/test_operators_identities.py:124-131); autograd path."""
    def make_nets():
        return [nd.FCNN(n_input_units=1, n_output_units=1, hidden_units=(16, 16))]

    def make_conditions():
        return [nd.IVP(t_0=0.0, u_0=1.0, u_0_prime=0.0)]

    def diff_eqs(u, t):
        return [nd.diff(u, t, order=3) + nd.diff(u, t)]

    return Workload("y1_third_order_ode", "Solver1D", ("t",), ((0.0, 2.0),), [((1, 16, 16, 1), "tanh")], make_nets,
                    make_conditions, diff_eqs, 1, 512, _fcnn_flops((1, 16, 16, 1), 4), None)


def _softplus_net(nd):
    """du/dt + u = 0 with a Softplus network: an activation without a jet rule in the kernels (like the reference's Swish /
    APTx, networks.py:155-208); autograd path."""
    def make_nets():
        return [nd.FCNN(n_input_units=1, n_output_units=1, hidden_units=(16, 16), actv=torch.nn.Softplus)]

    def make_conditions():
        return [nd.IVP(t_0=0.0, u_0=1.0)]

    def diff_eqs(u, t):
        return [nd.diff(u, t) + u]

    return Workload("y2_softplus_decay", "Solver1D", ("t",), ((0.0, 2.0),), [((1, 16, 16, 1), "softplus")], make_nets,
                    make_conditions, diff_eqs, 1, 512, _fcnn_flops((1, 16, 16, 1), 2), None)


def _biharmonic(nd):
    """Biharmonic plate  lap(lap u) = 1  on the unit square (DirichletBVP2D): NESTED operators -> fourth-order derivatives of the
    network output (the reference nests operators the same way, tests/test_operators_identities.py:124-131); autograd path."""
    def make_nets():
        return [nd.FCNN(n_input_units=2, n_output_units=1, hidden_units=(16, 16))]

    def make_conditions():
        return [nd.DirichletBVP2D(x_min=0, x_min_val=lambda y: 0, x_max=1, x_max_val=lambda y: 0,
                                  y_min=0, y_min_val=lambda x: 0, y_max=1, y_max_val=lambda x: 0)]

    def diff_eqs(u, x, y):
        return [nd.laplacian(nd.laplacian(u, x, y), x, y) - 1.0]

    return Workload("y3_biharmonic", "Solver2D", ("x", "y"), ((0.0, 1.0), (0.0, 1.0)), [((2, 16, 16, 1), "tanh")], make_nets,
                    make_conditions, diff_eqs, 1, 512, _fcnn_flops((2, 16, 16, 1), 15), None)


_EXTRA = {
    "x1": lambda nd: _heat(nd, "x1_heat_dirichlet_neumann", "right"),
    "x2": lambda nd: _heat(nd, "x2_heat_neumann_dirichlet", "left"),
    "x3": lambda nd: _bvp(nd, "x3_bvp_dirichlet_neumann", dict(x_min_val=1.0, x_max_prime=0.5)),
    "x4": lambda nd: _bvp(nd, "x4_bvp_neumann_dirichlet", dict(x_min_prime=-0.5, x_max_val=0.25)),
    "x5": lambda nd: _bvp(nd, "x5_bvp_neumann_neumann", dict(x_min_prime=-0.5, x_max_prime=0.5)),
    "x6": lambda nd: _bvp(nd, "x6_bvp_dirichlet_dirichlet", dict(x_min_val=1.0, x_max_val=0.25)),
    "x7": _ensemble,
    "x8": lambda nd: _heat(nd, "x8_heat_neumann_neumann", "both"),
    "x9": _resnet,
}
_BUILDERS = {"c1": _c1, "c2": _c2, "c3": _c3, "c4": _c4, "c5": _c5}
NAMES = tuple(_BUILDERS)          # BASELINE.json configs
EXTRA_NAMES = tuple(_EXTRA)       # widened condition family
_BUILDERS.update(_EXTRA)
# problems the fused engine refuses: they run on the autograd path (neurodiffeq_b200/eager.py)
_FALLBACK = {"y1": _third_order, "y2": _softplus_net, "y3": _biharmonic}
FALLBACK_NAMES = tuple(_FALLBACK)
_BUILDERS.update(_FALLBACK)


def build(nd, key):
    """``nd`` = namespace exposing the reference's public names; ``key`` in c1..c5."""
    return _BUILDERS[key](nd)


def sample_coords(workload, n, seed=0):
    """Synthetic uniform points in the workload's box: float32 array [d0, n] (SoA), deterministic per seed."""
    rs = np.random.RandomState(seed)
    rows = []
    for lo, hi in workload.coord_ranges:
        rows.append((lo + (hi - lo) * rs.rand(n)).astype(np.float32))
    return np.stack(rows, axis=0)


def bundle_eq_wrapper(workload):
    """diff_eqs as BundleSolver1D would call it (reference solvers.py:1353-1361): funcs, t, theta[eq_param_index]."""
    if workload.eq_param_index is None:
        return workload.diff_eqs
    n_funcs = len(workload.nets_spec) if workload.solver != "BundleSolver1D" else 2
    idx = tuple(n_funcs + 1 + i for i in workload.eq_param_index)

    def wrapped(*variables):
        head = variables[:n_funcs + 1]
        return workload.diff_eqs(*head, *(variables[i] for i in idx))

    return wrapped


# ----------------------------------------------------------------------------------------------------------------------
# operators.py golden cases: closed-form fields of three coordinates, usable on tensors and on traced symbols
# ----------------------------------------------------------------------------------------------------------------------
OPERATOR_NAMES = ("grad", "div", "curl", "laplacian", "vector_laplacian", "spherical_curl", "spherical_grad", "spherical_div",
                  "spherical_laplacian", "spherical_vector_laplacian", "spherical_to_cartesian", "cartesian_to_spherical",
                  "cylindrical_grad", "cylindrical_div", "cylindrical_curl", "cylindrical_laplacian",
                  "cylindrical_vector_laplacian", "cylindrical_to_cartesian", "cartesian_to_cylindrical")
_SCALAR_OPERATORS = ("grad", "laplacian", "spherical_grad", "spherical_laplacian", "cylindrical_grad", "cylindrical_laplacian")


def operator_fields(a, b, d):
    """three smooth closed-form fields of the coordinates (a, b, d)"""
    return (torch.sin(a) * b + d ** 2 * torch.cos(b), a * b * d + torch.exp(-a) * torch.sin(d), torch.cos(a * d) + b ** 2)


def operator_arguments(name, coords):
    """positional arguments of operator ``name`` for the golden case: (field(s)..., *coords) or just the coordinates"""
    if name.endswith("_to_cartesian") or name.startswith("cartesian_to"):
        return tuple(coords)
    f = operator_fields(*coords)
    return (f[0], *coords) if name in _SCALAR_OPERATORS else (*f, *coords)


# ---- the product on a workload (shared by tests/, bench.py and __graft_entry__.smoke()) --------------------------------------
def product_namespace():
    import types
    from neurodiffeq_b200 import diff
    from neurodiffeq_b200 import operators as ops
    from neurodiffeq_b200.networks import FCNN, SinActv, Resnet
    from neurodiffeq_b200 import conditions as c
    return types.SimpleNamespace(
        diff=diff, FCNN=FCNN, Resnet=Resnet, SinActv=SinActv, IVP=c.IVP, BundleIVP=c.BundleIVP, DirichletBVP2D=c.DirichletBVP2D,
        IBVP1D=c.IBVP1D, DirichletBVPSpherical=c.DirichletBVPSpherical, NoCondition=c.NoCondition,
        DoubleEndedBVP1D=c.DoubleEndedBVP1D, EnsembleCondition=c.EnsembleCondition,
        spherical_laplacian=ops.spherical_laplacian, laplacian=ops.laplacian, grad=ops.grad, div=ops.div,
        curl=ops.curl)


def distinct(nets):
    seen, out = set(), []
    for n in nets:
        if id(n) not in seen:
            seen.add(id(n))
            out.append(n)
    return out


def set_params(nets, arrays):
    import torch
    it = iter(arrays)
    with torch.no_grad():
        for m in distinct(nets):
            for p in m.parameters():
                p.copy_(torch.as_tensor(next(it), dtype=p.dtype).reshape(p.shape))


def build_fused(key, params=None, seed=0, device=None):
    """The product: trace the workload with neurodiffeq_b200's own classes and put it on the GPU."""
    import torch
    from neurodiffeq_b200.engine import FusedProblem
    wl = build(product_namespace(), key)
    torch.manual_seed(seed)
    nets, conds = wl.make_nets(), wl.make_conditions()
    if params is not None:
        set_params(nets, params)
    fp = FusedProblem(nets, conds, bundle_eq_wrapper(wl), len(wl.coord_names), device=device)
    return wl, nets, conds, fp
