"""CPU tests of the host-side mirror: generators, solver construction errors, program lowering details."""
import math

import numpy as np
import pytest
import torch

import workloads

from neurodiffeq_b200 import generators as G
from neurodiffeq_b200 import symbolic as S


def test_generator1d_methods_shapes_and_bounds():
    for method in ("uniform", "equally-spaced", "equally-spaced-noisy", "log-spaced", "log-spaced-noisy", "chebyshev",
                   "chebyshev1", "chebyshev2", "chebyshev2-noisy", "latin-hypercube"):
        g = G.Generator1D(64, t_min=0.1, t_max=2.0, method=method)
        x = g.get_examples()
        assert x.shape == (64,) and x.dtype == torch.float32 and not x.requires_grad
        if "noisy" not in method:
            assert x.min() >= 0.1 - 1e-6 and x.max() <= 2.0 + 1e-6
    with pytest.raises(ValueError):
        G.Generator1D(8, method="nope")
    with pytest.raises(ValueError):
        G.Generator1D(8, t_min=-1.0, t_max=1.0, method="log-spaced")
    g = G.Generator1D(1024, 0.1, 12.0, "equally-spaced-noisy")
    assert abs(g.noise_std - (11.9 / 1024) / 4) < 1e-12           # reference generators.py:149
    base = torch.linspace(0.1, 12.0, 1024)
    assert (g.get_examples() - base).abs().max() < 8 * g.noise_std


def test_generator2d_and_3d():
    g = G.Generator2D((16, 8), (0, -1), (1, 1), "equally-spaced")
    x, y = g.get_examples()
    assert g.size == 128 and x.shape == (128,) and y.shape == (128,)
    assert torch.allclose(x[:8], torch.zeros(8)) and torch.allclose(y[:8], torch.linspace(-1, 1, 8))   # 'ij' order
    gn = G.Generator2D((128, 128), (0, 0), (1, 1), "equally-spaced-noisy")
    xn, yn = gn.get_examples()
    xg, yg = G.Generator2D((128, 128), (0, 0), (1, 1), "equally-spaced").get_examples()
    assert (xn - xg).std() == pytest.approx(1 / 128 / 4, rel=0.1)   # N(0, (step/4)^2), generators.py:253-266
    g3 = G.Generator3D((4, 5, 6))
    assert g3.size == 120 and len(g3.get_examples()) == 3
    with pytest.raises(ValueError):
        G.Generator2D(method="bad")


def test_generator_spherical():
    g = G.GeneratorSpherical(4096, 0.1, 3.0)
    r, th, ph = g.get_examples()
    assert r.min() >= 0.1 and r.max() <= 3.0
    assert th.min() > 0 and th.max() < math.pi and ph.min() >= 0 and ph.max() <= 2 * math.pi + 1e-5
    assert torch.sin(th).min() > 1e-4              # never exactly on a pole
    assert abs((r ** 2).mean().item() - (0.01 + 9.0) / 2) < 0.2   # r^2 uniform
    with pytest.raises(ValueError):
        G.GeneratorSpherical(8, 2.0, 1.0)


def test_combinators():
    a, b = G.Generator1D(8, method="equally-spaced"), G.Generator1D(8, 2, 3, method="equally-spaced")
    assert (a + b).get_examples().shape == (16,) and (a + b).size == 16
    e = (a * b).get_examples()
    assert len(e) == 2 and e[0].shape == (8,) and (a * b).size == 8
    m = (a ^ b ^ G.Generator1D(4)).get_examples()
    assert len(m) == 3 and m[0].shape == (256,) and (a ^ b ^ G.Generator1D(4)).size == 256
    with pytest.raises(ValueError):
        G.EnsembleGenerator(a, G.Generator1D(9))
    st = G.StaticGenerator(G.Generator1D(8))
    assert torch.equal(st.get_examples(), st.get_examples())
    pre = G.PredefinedGenerator(np.arange(4.0), np.ones(4))
    assert [t.shape for t in G.SamplerGenerator(pre).get_examples()] == [(4, 1), (4, 1)]
    with pytest.raises(ValueError):
        a + 3


def test_register_allocation_is_small_and_interpreter_matches():
    """liveness-based slot reuse keeps the value file tiny; the numpy interpreter reproduces a closed form."""
    g = S.Graph()
    x, y = g.coord(0), g.coord(1)
    n = g.ych(0, 0, 0)
    expr = (x * y + torch.sin(3.0 * x)) / (1.0 + y * y) + n * torch.exp(-x)
    prog = S.lower([(S.OP_ST_U, 0, expr), (S.OP_ST_R, 0, expr * expr - 2.0)], lambda a, b, c: 0)
    assert prog.n_slots <= 6
    rs = np.random.RandomState(0)
    c = rs.rand(2, 50)
    yv = rs.rand(1, 50)
    u, r, _ = S.evaluate_program(prog, c, yv, n_u=1, n_r=1)
    want = (c[0] * c[1] + np.sin(3 * c[0])) / (1 + c[1] ** 2) + yv[0] * np.exp(-c[0])
    np.testing.assert_allclose(u[0], want, rtol=1e-12)
    np.testing.assert_allclose(r[0], want ** 2 - 2, rtol=1e-12)


def test_symbolic_reverse_matches_finite_differences():
    g = S.Graph()
    x = g.coord(0)
    n0, n1 = g.ych(0, 0, 0), g.ych(0, 0, 1)
    r = torch.tanh(n0 * x) + n1 ** 2 / (1.0 + x) - torch.cos(n0)
    adj = S.reverse_gradients([(r, g.const(1.0))])
    rows = {n0: 0, n1: 1}
    prog = S.lower([(S.OP_ST_R, 0, r)] + [(S.OP_ST_SEED, rows[k], v) for k, v in adj.items()],
                   lambda a, b, c: c)
    rs = np.random.RandomState(1)
    c, yv = rs.rand(1, 20) + 0.5, rs.rand(2, 20)
    _, r0, seed = S.evaluate_program(prog, c, yv, n_r=1, n_seed=2)
    for k in range(2):
        h = 1e-6
        yp, ym = yv.copy(), yv.copy()
        yp[k] += h
        ym[k] -= h
        fd = (S.evaluate_program(prog, c, yp, n_r=1, n_seed=2)[1] - S.evaluate_program(prog, c, ym, n_r=1, n_seed=2)[1]) / (2 * h)
        np.testing.assert_allclose(seed[k], fd[0], rtol=1e-6, atol=1e-8)


def test_unsupported_features_raise_loudly():
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200.networks import FCNN
    from neurodiffeq_b200.conditions import NoCondition, IBVP1D
    from neurodiffeq_b200 import diff
    import torch.nn as nn
    with pytest.raises(NotImplementedError):   # activation without a jet rule
        TracedProblem([FCNN(1, 1, actv=nn.ReLU)], [NoCondition()], lambda u, t: [diff(u, t)], 1)
    # Neumann IBVP evaluates the net at a boundary abscissa: one constant coordinate + a second instance of the module
    heat = lambda u, x, t: [diff(u, t) - diff(u, x, order=2)]  # noqa: E731
    net = FCNN(2, 1)
    tp = TracedProblem([net], [IBVP1D(0, 1, 0, lambda x: x, x_min_prime=lambda t: 0, x_max_val=lambda t: 0)], heat, 2)
    assert tp.const_coords == (0.0,) and tp.n_coords == 3 and [nd.in_coord for nd in tp.nets] == [(0, 1), (2, 1)]
    assert tp.nets[0].module is tp.nets[1].module
    # Neumann data on both ends: the two boundary abscissae never feed the same instance and share one jet direction
    both = TracedProblem([net], [IBVP1D(0, 1, 0, lambda x: x, x_min_prime=lambda t: 0, x_max_prime=lambda t: 0)], heat, 2)
    assert both.const_coords == (0.0, 1.0) and (both.scheme.n1, both.scheme.n2) == (4, 4)
    assert (0.0, 0.0, 1.0, 1.0) in both.scheme.dirs and (0.0, 1.0, 1.0, 1.0) in both.scheme.dirs
    with pytest.raises(NotImplementedError):   # all three mixed partials in 3-D: 6 jet directions, the kernels carry 4
        from neurodiffeq_b200.engine import pad_scheme
        TracedProblem([FCNN(3, 1)], [NoCondition()],
                      lambda u, x, y, z: [diff(diff(u, x), y) + diff(diff(u, x), z) + diff(diff(u, y), z)], 3,
                      pad_scheme=pad_scheme)
    with pytest.raises(RuntimeError):          # a boundary leaf outside a trace has no meaning
        S.Graph().const_coord(1.0)
    with pytest.raises(NotImplementedError):   # torch.cat of traced columns outside enforce()
        TracedProblem([FCNN(1, 1)], [NoCondition()], lambda u, t: [torch.cat([u, t], 1)], 1)


@pytest.mark.parametrize("key", workloads.NAMES + workloads.EXTRA_NAMES)
def test_conditions_and_diff_on_eager_tensors_equal_the_oracle(key):
    """Outside the solver the product's conditions / diff / operators are ordinary torch code with the reference's
    semantics (conditions.py:41-57 and the parameterize of each class): same u and residual as the oracle, incl. the
    Neumann branches that evaluate the network at a boundary leaf."""
    from helpers import product_namespace, get_params
    from oracle import reference_port as oracle
    wl = workloads.build(product_namespace(), key)
    torch.manual_seed(1)
    nets, conds = wl.make_nets(), wl.make_conditions()
    for m in nets:
        m.double()
    owl = workloads.build(oracle.NAMESPACE, key)
    onets, oconds = owl.make_nets(), owl.make_conditions()
    oracle.load_params(onets, get_params(nets), dtype=torch.float64)
    coords_np = workloads.sample_coords(wl, 64, seed=3)

    def run(nets_, conds_, w):
        cols = [torch.as_tensor(c, dtype=torch.float64).reshape(-1, 1).requires_grad_(True) for c in coords_np]
        funcs = [c.enforce(n, *cols) for n, c in zip(nets_, conds_)]
        res = workloads.bundle_eq_wrapper(w)(*funcs, *cols)
        return [f.detach().numpy() for f in funcs], [r.detach().numpy() for r in res]

    fu, fr = run(nets, conds, wl)
    gu, gr = run(onets, oconds, owl)
    for a, b in zip(fu + fr, gu + gr):
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-12)


def test_wrapping_generators():
    """TransformGenerator / FilterGenerator / ResampleGenerator / BatchGenerator / GeneratorND
    (reference generators.py:419-570, 758-801, 904-1043)."""
    base = G.Generator1D(64, 0.0, 1.0, "equally-spaced") * G.Generator1D(64, 2.0, 3.0, "equally-spaced")
    x, y = G.TransformGenerator(base, transforms=[lambda a: 2 * a, None]).get_examples()
    assert torch.allclose(x, 2 * torch.linspace(0, 1, 64)) and torch.allclose(y, torch.linspace(2, 3, 64))
    s, d = G.TransformGenerator(base, transform=lambda a, b: (a + b, a - b)).get_examples()
    assert torch.allclose(s - d, 2 * torch.linspace(2, 3, 64))
    with pytest.raises(ValueError):
        G.TransformGenerator(base, transforms=[None, None], transform=lambda a, b: (a, b))
    f = G.FilterGenerator(base, lambda xs: xs[0] > 0.5)
    fx, fy = f.get_examples()
    assert f.size == len(fx) == len(fy) == 32 and fx.min() > 0.5
    r = G.ResampleGenerator(base, size=10)
    rx, ry = r.get_examples()
    assert len(rx) == 10 and torch.allclose(ry - rx, torch.full((10,), 2.0)) and len(set(rx.tolist())) == 10
    rb = G.ResampleGenerator(G.Generator1D(8, 0, 1, "equally-spaced"), size=50, replacement=True).get_examples()
    assert rb.shape == (50,)
    b = G.BatchGenerator(G.Generator1D(10, 0.0, 1.0, "equally-spaced"), 4)
    got = torch.cat([b.get_examples() for _ in range(5)])       # 20 samples = two passes over the 10 cached nodes
    assert b.size == 4 and torch.allclose(got, torch.linspace(0, 1, 10).repeat(2))
    with pytest.raises(ValueError):
        G.BatchGenerator(G.FilterGenerator(base, lambda xs: xs[0] > 2, size=0, update_size=False), 4)
    nd = G.GeneratorND(grid=(4, 5, 3), r_min=(0.0, 1.0, 0.0), r_max=(1.0, 100.0, 2.0),
                       methods=["equally-spaced", "log-spaced", "exp-spaced"], noisy=False)
    a, bb, c = nd.get_examples()
    assert nd.size == 60 and a.shape == bb.shape == c.shape == (60,)
    assert torch.allclose(bb.reshape(4, 5, 3)[0, :, 0], torch.logspace(0, 2, 5))
    assert torch.allclose(10 ** c.reshape(4, 5, 3)[0, 0], torch.linspace(1.0, 100.0, 3), rtol=1e-5)
    noisy = G.GeneratorND(grid=(16,), r_min=0.0, r_max=1.0, methods="chebyshev2", noisy=True, abs_value=True, cut=(2, -2))
    pts, = noisy.get_examples()
    assert pts.shape == (12,) and (pts >= 0).all()
    with pytest.raises(ValueError):
        G.GeneratorND(grid=(4,), r_min=0.0, r_max=1.0, methods="sobol")
    with pytest.raises(ValueError):
        G.GeneratorND(grid=(4,), r_min=0.0, r_max=1.0, methods="uniform", stride=2)


def test_deprecated_argument_names_behave_like_the_reference():
    """reference tests/test_conditions.py:235-350 and _version_utils.py:21-48: old keyword names work with a
    FutureWarning, both spellings at once are a KeyError, unknown bundle parameters a ValueError."""
    from neurodiffeq_b200.conditions import IVP, BundleIVP, DirichletBVP, BundleDirichletBVP
    from neurodiffeq_b200 import diff
    with pytest.warns(FutureWarning):
        c = IVP(0, x_0=1)
    assert c.u_0 == 1
    with pytest.warns(FutureWarning):
        assert IVP(0, 1, x_0_prime=2).u_0_prime == 2
    with pytest.raises(KeyError):
        IVP(0, x_0=1, u_0=2)
    with pytest.raises(KeyError):
        IVP(0, x_0_prime=1, u_0_prime=2)
    for bad in ("magic", "t0", "u1", "u1prime"):
        with pytest.raises(ValueError):
            BundleIVP(0.0, 1.0, 2.0, bundle_param_lookup={bad: 0})
    for bad in ("magic", "u0", "t0", "u1", "t1"):
        with pytest.raises(ValueError):
            BundleDirichletBVP(0.0, 1.0, 2.0, 3.0, bundle_param_lookup={bad: 0})
    with pytest.warns(FutureWarning):
        BundleIVP(0.0, 1.0, bundle_conditions={"t_0": 0})
    with pytest.warns(FutureWarning):
        BundleDirichletBVP(0.0, 1.0, 2.0, 3.0, bundle_conditions={"t_0": 0})
    with pytest.warns(FutureWarning):
        b = DirichletBVP(t_0=0, t_1=1, x_0=2, x_1=3)
    assert (b.u_0, b.u_1) == (2, 3)
    with pytest.raises(KeyError):
        DirichletBVP(t_0=0, u_0=0, x_0=0, t_1=0, x_1=0)
    with pytest.raises(KeyError), pytest.warns(FutureWarning):       # x_0 renamed (warning), then x_1 clashes with u_1
        DirichletBVP(t_0=0, x_0=0, t_1=0, x_1=0, u_1=0)
    t = torch.linspace(0, 1, 5).reshape(-1, 1).requires_grad_(True)
    with pytest.warns(FutureWarning):
        assert torch.allclose(diff(x=t ** 2, t=t), 2 * t)


def test_network_modules_of_the_reference():
    """Resnet / MonomialNN / Swish / APTx (reference networks.py:73-208, tests/test_networks.py): eager semantics; the fused
    path runs Resnet (body on the kernels, shortcut in the program) and refuses the others with a clear NotImplementedError."""
    import torch.nn as nn
    from neurodiffeq_b200.networks import FCNN, Resnet, MonomialNN, Swish, APTx
    from neurodiffeq_b200.conditions import NoCondition
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200 import diff
    x = torch.linspace(-1, 1, 7).reshape(-1, 1)
    r = Resnet(1, 2, hidden_units=(8,))
    assert torch.allclose(r(x), r.skip_connection(x) + r.residual(x)) and r(x).shape == (7, 2)
    assert r.skip_connection.bias is None
    m = MonomialNN(3)
    assert m.degrees == (1, 2, 3) and torch.allclose(m(x), torch.cat([x, x ** 2, x ** 3], 1))
    assert MonomialNN([2, 4])(x).shape == (7, 2) and "degrees=(2, 4)" in repr(MonomialNN([2, 4]))
    with pytest.raises(ValueError):
        MonomialNN([])
    with pytest.warns(UserWarning):
        MonomialNN([0, 1])
    with pytest.warns(UserWarning):
        MonomialNN([1, 1])
    assert torch.allclose(Swish(2.0)(x), x * torch.sigmoid(2.0 * x))
    assert len(list(Swish(trainable=True).parameters())) == 1 and len(list(Swish().parameters())) == 0
    assert torch.allclose(APTx(1.0, 2.0, 0.5)(x), (1.0 + torch.tanh(2.0 * x)) * 0.5 * x)
    assert len(list(APTx(trainable=True).parameters())) == 3
    tp = TracedProblem([Resnet(1, 1)], [NoCondition()], lambda u, t: [diff(u, t)], 1)     # Resnet: body + program scalars
    assert tp.nets[0].skip is not None and len(tp.prog_eval.patch) == 1
    for net in (FCNN(1, 1, actv=Swish), nn.Sequential(MonomialNN(2), nn.Linear(2, 1))):
        with pytest.raises(NotImplementedError):
            TracedProblem([net], [NoCondition()], lambda u, t: [diff(u, t)], 1)


def test_engine_helpers_for_resnet_shortcuts_on_the_cpu():
    """The two pieces of engine logic a Resnet adds -- re-writing program immediates from the parameters and the shortcut
    gradient from the seeds in the workspace -- are plain tensor code: exercised here on CPU tensors with a hand-laid-out
    workspace (the layout K1 writes: [tile][row][point]) against the float64 mirror."""
    from helpers import product_namespace, get_params
    from neurodiffeq_b200.engine import FusedProblem, pad_scheme, combine_seconds
    from neurodiffeq_b200.tracing import TracedProblem
    from oracle import jet_numpy
    wl = workloads.build(product_namespace(), "x9")
    torch.manual_seed(5)
    nets, conds = wl.make_nets(), wl.make_conditions()
    tp = TracedProblem(nets, conds, wl.diff_eqs, 2, pad_scheme=pad_scheme, combine_seconds=combine_seconds)
    fp = object.__new__(FusedProblem)                    # no CUDA library: only the pure-tensor helpers are used
    fp.device, fp.tp = torch.device("cpu"), tp
    fp._adopt_parameters()
    fp._register_program_scalars()
    n, T = 37, 16
    coords = workloads.sample_coords(wl, n, seed=2)
    params = get_params(nets)
    ref = jet_numpy.run_traced(tp, [params], coords)
    # (1) immediates follow the parameters
    dev_prog = fp._upload(tp.prog_train)
    pcs = sorted(tp.prog_train.patch)
    assert pcs and all(dev_prog[pc, 0] == S.OP_CONST for pc in pcs)
    w_skip = nets[0].skip_connection.weight.detach().float().reshape(-1)
    got = dev_prog[pcs, 2].view(torch.float32)
    want = torch.stack([w_skip[tp.prog_train.patch[pc][3] + 2 * tp.prog_train.patch[pc][2]] for pc in pcs])
    assert torch.equal(got, want)
    with torch.no_grad():
        nets[0].skip_connection.weight.mul_(2.0)
    fp._apply_patches()
    assert torch.equal(dev_prog[pcs, 2].view(torch.float32), 2 * want)
    # (2) shortcut gradient from the seeds in a workspace laid out like K1's
    nt = (n + T - 1) // T
    ws_seed = 256
    seeds = np.zeros((nt, tp.n_yrows, T), dtype=np.float32)
    for p in range(n):
        seeds[p // T, :, p % T] = ref["seeds"][:, p]
    fp.workspace = torch.zeros(ws_seed + seeds.nbytes, dtype=torch.uint8)
    fp.workspace[ws_seed:] = torch.from_numpy(seeds.reshape(-1).view(np.uint8))
    fp._plan_cache = {n: dict(T=T, n_tiles=nt, ws_seed=ws_seed)}
    fp.grad.zero_()
    fp._accumulate_shortcut_grads([torch.from_numpy(c) for c in coords], n)
    base = fp._skips[0][1]
    np.testing.assert_allclose(fp.grad[base: base + 2].numpy(), ref["grads"][-1].reshape(-1), rtol=2e-5)
