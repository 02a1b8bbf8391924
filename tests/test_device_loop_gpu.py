"""GPU: sampling on the device (pj_sample) and the opt-in device loop of the solvers (one CUDA-graph replay
per epoch: sampling + K0..K2b + best-parameter bookkeeping + FlatAdam).  Fixed-node generators must reproduce the host
generators exactly; random ones follow the same law (reference generators.py:107-191, 194-314, 572-655), checked by moments;
with fixed-node generators the device loop must walk the same trajectory as the host loop."""
import math

import numpy as np
import pytest
import torch

from helpers import product_namespace
import workloads

pytestmark = pytest.mark.gpu


def _host(gen):
    ex = gen.get_examples()
    return [t.numpy() for t in ((ex,) if isinstance(ex, torch.Tensor) else ex)]


def _dev(sampler):
    ex = sampler.get_examples()
    return [t.cpu().numpy() for t in ((ex,) if isinstance(ex, torch.Tensor) else ex)]


@pytest.mark.parametrize("method", ["equally-spaced", "log-spaced", "chebyshev1", "chebyshev2"])
def test_fixed_node_generators_are_reproduced_exactly(method):
    from neurodiffeq_b200 import generators as G
    from neurodiffeq_b200.device_sampling import DeviceSampler
    g = G.Generator1D(1000, 0.5, 3.0, method=method)
    np.testing.assert_array_equal(_dev(DeviceSampler(g, "cuda"))[0], _host(g)[0])
    g2 = G.Generator2D((17, 23), (0.0, -1.0), (2.0, 1.0), method="equally-spaced")
    for a, b in zip(_dev(DeviceSampler(g2, "cuda")), _host(g2)):
        np.testing.assert_array_equal(a, b)
    mesh = G.Generator1D(7, 0.0, 1.0, method="equally-spaced") ^ G.Generator1D(5, 2.0, 3.0, method="chebyshev1")
    for a, b in zip(_dev(DeviceSampler(mesh, "cuda")), _host(mesh)):
        np.testing.assert_array_equal(a, b)
    pre = G.PredefinedGenerator(torch.linspace(0, 1, 33), torch.linspace(5, 6, 33))
    for a, b in zip(_dev(DeviceSampler(pre, "cuda")), _host(pre)):
        np.testing.assert_array_equal(a, b)


def test_random_laws_have_the_reference_moments():
    from neurodiffeq_b200 import generators as G
    from neurodiffeq_b200.device_sampling import DeviceSampler
    n = 200000
    u = _dev(DeviceSampler(G.Generator1D(n, -2.0, 3.0, method="uniform"), "cuda", seed=1))[0]
    assert u.min() >= -2.0 and u.max() < 3.0
    assert abs(u.mean() - 0.5) < 5 * (5.0 / math.sqrt(12 * n)) and abs(u.var() - 25.0 / 12) < 0.02
    g = G.Generator1D(n, 0.0, 1.0, method="equally-spaced-noisy")           # N(0, (step / 4)^2) around the grid nodes
    d = _dev(DeviceSampler(g, "cuda", seed=2))[0] - np.linspace(0.0, 1.0, n, dtype=np.float32)
    assert abs(d.mean()) < 5 * g.noise_std / math.sqrt(n) and abs(d.std() / g.noise_std - 1.0) < 0.01
    assert abs(np.mean(d ** 4) / g.noise_std ** 4 - 3.0) < 0.1                # Gaussian kurtosis
    g2 = G.Generator2D((400, 500), (0.0, 0.0), (1.0, 2.0))                   # default: equally-spaced-noisy
    xs, ys = _dev(DeviceSampler(g2, "cuda", seed=3))
    bx, by = [p.numpy() for p in g2._static]
    assert abs((xs - bx).std() / g2._std[0] - 1.0) < 0.01 and abs((ys - by).std() / g2._std[1] - 1.0) < 0.01
    assert abs(np.corrcoef(xs - bx, ys - by)[0, 1]) < 0.01                   # independent jitter per coordinate
    sph = G.GeneratorSpherical(n, 0.5, 2.0)
    r, th, ph = _dev(DeviceSampler(sph, "cuda", seed=4))
    rh, thh, phh = _host(sph)
    assert r.min() >= 0.5 and r.max() <= 2.0 and th.min() > 0 and th.max() < math.pi and ph.min() >= 0 and ph.max() <= 2 * math.pi
    for a, b in ((r, rh), (th, thh), (ph, phh), (np.cos(th), np.cos(thh))):  # same law as the host generator
        assert abs(a.mean() - b.mean()) < 0.01 and abs(a.std() - b.std()) < 0.01
    mesh = G.Generator1D(300, 0.0, 1.0, method="equally-spaced-noisy") ^ G.Generator1D(200, 0.0, 1.0, method="uniform")
    mx, my = [a.reshape(300, 200) for a in _dev(DeviceSampler(mesh, "cuda", seed=5))]
    assert np.all(mx == mx[:, :1]) and np.all(my == my[:1, :])               # one draw per NODE, shared along the other axis


def test_calls_differ_and_rows_are_a_function_of_the_global_index():
    from neurodiffeq_b200 import generators as G
    from neurodiffeq_b200.device_sampling import DeviceSampler
    g = G.Generator2D((64, 64), (0.0, 0.0), (1.0, 1.0))
    a, b = DeviceSampler(g, "cuda", seed=9), DeviceSampler(g, "cuda", seed=9)
    full1, full2 = a.get_examples(), a.get_examples()
    assert not torch.equal(full1[0], full2[0])                               # a new call draws new points
    lo, hi = 1000, 3077                                                      # a rank's shard: rows [lo, hi) of the same batch
    part = [torch.empty(hi - lo, device="cuda") for _ in range(2)]
    b.sample_into(part, lo, hi - lo)
    for p, f in zip(part, full1):
        assert torch.equal(p, f[lo:hi])
    b.sample_into(part, lo, hi - lo)
    for p, f in zip(part, full2):
        assert torch.equal(p, f[lo:hi])


def _laplace_solver(device_loop, method, n_valid, seed=0):
    from neurodiffeq_b200 import solvers as S, generators as G
    from neurodiffeq_b200.optim import FlatAdam
    wl = workloads.build(product_namespace(), "c2")
    torch.manual_seed(seed)
    nets, conds = wl.make_nets(), wl.make_conditions()
    tg = G.Generator2D((48, 48), (0.0, 0.0), (1.0, 1.0), method=method)
    vg = G.Generator2D((24, 24), (0.0, 0.0), (1.0, 1.0), method="equally-spaced")
    solver = S.Solver2D(wl.diff_eqs, conds, nets=nets, train_generator=tg, valid_generator=vg, n_batches_valid=n_valid,
                        device_loop=device_loop)
    if not device_loop:
        solver.optimizer = FlatAdam.for_solver(solver)
    return solver


@pytest.mark.parametrize("n_valid", [0, 2])
def test_device_loop_walks_the_host_loop_trajectory(n_valid):
    host, dev = _laplace_solver(False, "equally-spaced", n_valid), _laplace_solver(True, "equally-spaced", n_valid)
    assert dev._device_loop_blocker() is None
    host.fit(25, tqdm_file=None)
    dev.fit(10, tqdm_file=None)
    dev.fit(15, tqdm_file=None)                                               # a second fit() continues the same run
    for key in ("train_loss",) + (("valid_loss",) if n_valid else ()):
        a, b = np.array(host.metrics_history[key]), np.array(dev.metrics_history[key])
        assert len(a) == len(b) == 25
        np.testing.assert_allclose(b, a, rtol=2e-4)
    assert abs(dev.lowest_loss - host.lowest_loss) <= 2e-4 * host.lowest_loss
    x = torch.linspace(0.05, 0.95, 50)
    ua, ub = host.get_solution(best=True)(x, x, to_numpy=True), dev.get_solution(best=True)(x, x, to_numpy=True)
    np.testing.assert_allclose(ub, ua, rtol=1e-3, atol=1e-5)


def test_device_loop_with_noisy_sampling_and_callbacks():
    solver = _laplace_solver(True, "equally-spaced-noisy", 1)
    seen = []
    solver.fit(60, callbacks=[lambda s: seen.append((s.local_epoch, len(s.metrics_history["train_loss"]), s.lowest_loss))],
               tqdm_file=None)
    assert [e for e, _, _ in seen] == list(range(1, 61)) and [k for _, k, _ in seen] == list(range(1, 61))
    assert all(low is not None for _, _, low in seen)
    tl = solver.metrics_history["train_loss"]
    assert len(set(tl)) > 50 and np.mean(tl[-10:]) < 0.5 * np.mean(tl[:10])   # fresh points every epoch, and it trains
    assert solver.lowest_loss == pytest.approx(min(solver.metrics_history["valid_loss"]), rel=1e-6)


def test_device_loop_falls_back_with_a_reason():
    from neurodiffeq_b200 import solvers as S, generators as G
    wl = workloads.build(product_namespace(), "c2")
    nets, conds = wl.make_nets(), wl.make_conditions()
    tg = G.Generator2D((16, 16), (0.0, 0.0), (1.0, 1.0), method="chebyshev2-noisy")   # redrawn axes: no device law
    solver = S.Solver2D(wl.diff_eqs, conds, nets=nets, train_generator=tg, valid_generator=tg, n_batches_valid=0, device_loop=True)
    with pytest.warns(RuntimeWarning, match="device_loop=True is not possible"):
        solver.fit(3, tqdm_file=None)
    assert len(solver.metrics_history["train_loss"]) == 3
