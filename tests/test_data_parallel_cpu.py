"""N>1 host logic on CPU: two gloo ranks shard a batch, compute partial gradients with the GLOBAL loss scale (numpy
mirror of the kernels = test infrastructure), all-reduce the flat [grad | sum r^2] buffer and must reproduce the
single-process result of the reference (golden vectors)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, key, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import workloads
    from conftest import load_golden
    from helpers import product_namespace
    from neurodiffeq_b200.tracing import TracedProblem
    from neurodiffeq_b200.parallel import shard_bounds, all_reduce_gradbuf
    from oracle import jet_numpy
    wl = workloads.build(product_namespace(), key)
    gold = load_golden(wl.name)
    nets, conds = wl.make_nets(), wl.make_conditions()
    tp = TracedProblem(nets, conds, workloads.bundle_eq_wrapper(wl), len(wl.coord_names))
    per_net, it = [], iter(gold["params"])
    for nd in tp.nets:
        per_net.append([next(it) for _ in range(2 * len(nd.linears))])
    n = gold["coords"].shape[1]
    lo, hi = shard_bounds(n, rank, world)
    out = jet_numpy.run_traced(tp, per_net, gold["coords"][:, lo:hi], n_global=n)
    flat = np.concatenate([g.reshape(-1) for g in out["grads"]] + [np.array([(out["residual"] ** 2).sum()])])
    buf = torch.from_numpy(flat)
    all_reduce_gradbuf(buf, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "buf.npy"), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("key", ["c2", "c5"])
def test_two_rank_sharding_reproduces_single_process_gradient(key, tmp_path):
    from conftest import load_golden
    import workloads
    from helpers import product_namespace
    world = 2
    port = 29000 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, key, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    buf = np.load(os.path.join(str(tmp_path), "buf.npy"))
    wl = workloads.build(product_namespace(), key)
    gold = load_golden(wl.name)
    flat = np.concatenate([g.reshape(-1) for g in gold["grads"]])
    n = gold["coords"].shape[1]
    assert np.linalg.norm(buf[:-1] - flat) <= 1e-9 * np.linalg.norm(flat)
    assert abs(buf[-1] / (n * gold["residual"].shape[0]) - gold["loss"]) <= 1e-10 * gold["loss"]


def test_shard_bounds_tile_the_batch():
    from neurodiffeq_b200.parallel import shard_bounds
    for n in (1, 7, 16384, 65537):
        for w in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def _mean_u(u, x, y):
    return u.mean()


def _funcs_loss(residual, funcs, coords):     # module level: picklable for the spawned ranks
    return residual.abs().mean() + 0.2 * (funcs[0] ** 2).mean()


def _solver_worker(rank, world, port, key, out_dir, loss_fn=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    import neurodiffeq_b200.solvers as S
    from cpu_engine import CpuFusedProblem
    from test_solvers_gpu import make_solver
    S.FusedProblem = CpuFusedProblem                    # the stand-in engine; the data-parallel logic is the product's
    kw = {} if loss_fn is None else dict(loss_fn=loss_fn)
    if key == "c2":
        kw["metrics"] = {"mean_u": _mean_u}
    if key.startswith("y"):        # refused by the tracer: the autograd fallback (eager.py), here on CPU tensors
        kw["device"] = "cpu"
    n_pts = 151 if key == "x4" or kw.get("loss_fn") == "l1" else 150      # odd: the two ranks get 76 / 75 points
    wl, solver, nets, coords_np = make_solver(key, n_pts, **kw)  # same seed on every rank -> same parameters, same batch
    assert solver._dist is not None
    solver.fit(3, tqdm_file=None)
    from helpers import get_params
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), train=np.array(solver.metrics_history["train_loss"]),
             valid=np.array(solver.metrics_history["valid_loss"]),
             metric=np.array(solver.metrics_history.get("train__mean_u", [])), theta=np.concatenate([p.reshape(-1) for p in get_params(nets)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("key,loss_fn", [("c2", None), ("x4", None), ("c2", "l1"), ("c2", _funcs_loss), ("y1", None)])
def test_solver_fit_on_two_ranks_equals_one_process(key, loss_fn, tmp_path, monkeypatch):
    """Solver.fit under torch.distributed: every rank samples the same batch, keeps its slice, the flat
    [grad | sum r^2] buffer is all-reduced once per epoch phase and the replicated Adam stays in lock-step -- losses and
    parameters equal the single-process run, and the ranks agree bit for bit."""
    world = 2
    port = 31000 + (os.getpid() % 2000)
    mp.start_processes(_solver_worker, args=(world, port, key, str(tmp_path), loss_fn), nprocs=world, join=True,
                       start_method="spawn")
    r0, r1 = np.load(os.path.join(str(tmp_path), "rank0.npz")), np.load(os.path.join(str(tmp_path), "rank1.npz"))
    assert np.array_equal(r0["theta"], r1["theta"]) and np.array_equal(r0["train"], r1["train"])
    assert np.array_equal(r0["metric"], r1["metric"]) and (key != "c2" or len(r0["metric"]) == 3)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import neurodiffeq_b200.solvers as S
    from cpu_engine import CpuFusedProblem
    from test_solvers_gpu import make_solver
    from helpers import get_params
    monkeypatch.setattr(S, "FusedProblem", CpuFusedProblem)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        n_pts = 151 if key == "x4" or loss_fn == "l1" else 150
        extra = dict(device="cpu") if key.startswith("y") else {}
        wl, solver, nets, _ = make_solver(key, n_pts, **({} if loss_fn is None else dict(loss_fn=loss_fn)), **extra)
        solver.fit(3, tqdm_file=None)
    finally:
        torch.set_default_dtype(old)
    np.testing.assert_allclose(r0["train"], solver.metrics_history["train_loss"], rtol=1e-6)
    np.testing.assert_allclose(r0["valid"], solver.metrics_history["valid_loss"], rtol=1e-6)
    theta = np.concatenate([p.reshape(-1) for p in get_params(nets)])
    np.testing.assert_allclose(r0["theta"], theta, rtol=1e-8, atol=1e-11)


def _lbfgs_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    import workloads
    import neurodiffeq_b200.solvers as S
    from cpu_engine import CpuFusedProblem
    from helpers import product_namespace, get_params
    from neurodiffeq_b200.generators import PredefinedGenerator
    S.FusedProblem = CpuFusedProblem
    wl = workloads.build(product_namespace(), "x6")
    torch.manual_seed(0)
    nets = wl.make_nets()
    gen = PredefinedGenerator(*workloads.sample_coords(wl, 90, seed=21))
    opt = torch.optim.LBFGS([p for m in nets for p in m.parameters()], lr=0.5, max_iter=3, history_size=4)
    solver = S.Solver1D(wl.diff_eqs, wl.make_conditions(), nets=nets, train_generator=gen, valid_generator=gen,
                        n_batches_valid=1, optimizer=opt)
    solver.fit(2, tqdm_file=None)
    np.savez(os.path.join(out_dir, f"lbfgs_w{world}_r{rank}.npz"), train=np.array(solver.metrics_history["train_loss"]),
             theta=np.concatenate([p.reshape(-1) for p in get_params(nets)]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_lbfgs_closure_mode_on_two_ranks(tmp_path):
    """Every closure evaluation all-reduces [grad | loss]; LBFGS then takes identical decisions on every rank."""
    port = 33000 + (os.getpid() % 2000)
    mp.start_processes(_lbfgs_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    mp.start_processes(_lbfgs_worker, args=(1, port + 1, str(tmp_path)), nprocs=1, join=True, start_method="spawn")
    a, b = np.load(os.path.join(str(tmp_path), "lbfgs_w2_r0.npz")), np.load(os.path.join(str(tmp_path), "lbfgs_w2_r1.npz"))
    one = np.load(os.path.join(str(tmp_path), "lbfgs_w1_r0.npz"))
    assert np.array_equal(a["theta"], b["theta"])
    np.testing.assert_allclose(a["train"], one["train"], rtol=1e-6)
    np.testing.assert_allclose(a["theta"], one["theta"], rtol=1e-6, atol=1e-9)
