"""torchrun worker of tests/test_round2_gpu.py::test_two_nccl_ranks_reproduce_the_single_gpu_gradient (one rank per GPU).

  python -m torch.distributed.run --nproc-per-node 2 ... tests/dp_nccl_worker.py <workload> <out.json>

Every rank (a) evaluates the WHOLE batch alone, (b) evaluates its slice with the global loss scale and all-reduces the flat
[grad | sum r^2] buffer over NCCL, (c) runs Solver.fit under data parallelism; rank 0 writes the deviations."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    key, out = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import workloads
    from helpers import build_fused, get_params
    from neurodiffeq_b200.parallel import shard_bounds, all_reduce_gradbuf, GradBufReducer
    from test_solvers_gpu import make_solver

    wl, nets, conds, fp = build_fused(key, seed=4)              # same seed: replicated parameters
    n = 6001                                                    # odd: unequal shards
    coords = [torch.from_numpy(c).cuda() for c in workloads.sample_coords(wl, n, seed=8)]
    fp.gradbuf.zero_()
    fp.residual_grad(coords, n_global=n, sumsq_out=fp.sumsq)
    full = fp.gradbuf.clone()
    lo, hi = shard_bounds(n, rank, world)
    fp.gradbuf.zero_()
    fp.residual_grad([c[lo:hi].contiguous() for c in coords], n_global=n, sumsq_out=fp.sumsq)
    shard = fp.gradbuf.clone()
    all_reduce_gradbuf(fp.gradbuf, dist)
    torch.cuda.synchronize()
    nccl_sum = fp.gradbuf.clone()
    grad_rel = float((fp.gradbuf[:-1] - full[:-1]).norm() / full[:-1].norm())
    sumsq_rel = float(abs(fp.gradbuf[-1] - full[-1]) / full[-1])
    # the hand-written one-shot NVLink all-reduce: same sum as NCCL's (rank order vs. NCCL's order: fp32 rounding), identical
    # on every rank, repeatable back to back (epoch / double-buffer protocol) and inside a replayed CUDA graph
    red = GradBufReducer(shard, dist)
    oneshot = {"mode": red.mode, "why": red.why}
    if red.mode == "oneshot-nvlink":
        work = shard.clone()
        red2 = red
        errs, same = [], True
        for it in range(7):
            work.copy_(shard * float(it + 1))
            red2(work)
            torch.cuda.synchronize()
            ref = nccl_sum * float(it + 1)
            errs.append(float((work - ref).norm() / ref.norm()))
            g = [torch.empty_like(work) for _ in range(world)]
            dist.all_gather(g, work)
            same = same and all(bool(torch.equal(x, g[0])) for x in g)
        graph = torch.cuda.CUDAGraph()
        static = shard.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            red2(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.graph(graph):
            red2(static)
        for it in range(3):
            static.copy_(shard * float(it + 2))
            graph.replay()
            torch.cuda.synchronize()
            ref = nccl_sum * float(it + 2)
            errs.append(float((static - ref).norm() / ref.norm()))
        oneshot.update(max_rel_err_vs_nccl=max(errs), ranks_identical=same)
        # K2b + collective as ONE kernel (pj_backward_allreduce): same arithmetic in the same order as K2b followed by the
        # one-shot kernel -> bit-identical to it; back to back, accumulating into a non-zero buffer, and in a replayed graph
        oneshot["fused"] = red.fused_args is not None
        if red.fused_args is not None:
            shard_coords = [c[lo:hi].contiguous() for c in coords]
            two_step = shard.clone()
            red(two_step)
            torch.cuda.synchronize()
            fused_equal, fused_same, fused_rel = True, True, 0.0
            for it in range(5):
                fp.gradbuf.zero_()
                fp.residual_grad(shard_coords, n_global=n, sumsq_out=fp.sumsq, reducer=red)
                torch.cuda.synchronize()
                fused_equal = fused_equal and bool(torch.equal(fp.gradbuf, two_step))
                fused_rel = max(fused_rel, float((fp.gradbuf - two_step).norm() / two_step.norm()))
                g = [torch.empty_like(fp.gradbuf) for _ in range(world)]
                dist.all_gather(g, fp.gradbuf)
                fused_same = fused_same and all(bool(torch.equal(x, g[0])) for x in g)
            fp.gradbuf.fill_(0.25)                      # accumulation: every rank's previous content is summed along
            fp.residual_grad(shard_coords, n_global=n, sumsq_out=fp.sumsq, reducer=red)
            torch.cuda.synchronize()
            acc_ref = two_step + 0.25 * world
            fused_acc = float((fp.gradbuf - acc_ref).norm() / acc_ref.norm())
            graph = torch.cuda.CUDAGraph()

            def fused_body():
                fp.gradbuf.zero_()
                fp.residual_grad(shard_coords, n_global=n, sumsq_out=fp.sumsq, reducer=red)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fused_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            dist.barrier()
            with torch.cuda.graph(graph):
                fused_body()
            for it in range(3):
                graph.replay()
                torch.cuda.synchronize()
                fused_equal = fused_equal and bool(torch.equal(fp.gradbuf, two_step))
            oneshot.update(fused_rel_vs_two_step=fused_rel, fused_equals_two_step=fused_equal, fused_ranks_identical=fused_same, fused_accumulate_rel=fused_acc)

    # Solver.fit in lock-step: the ranks end with identical parameters, equal to a single-process run of the same problem
    wl, solver, snets, coords_np = make_solver(key, 3001)
    assert solver._dist is not None
    solver.fit(4, tqdm_file=None)
    theta = torch.cat([torch.as_tensor(p).reshape(-1) for p in get_params(snets)]).cuda()
    gathered = [torch.empty_like(theta) for _ in range(world)]
    dist.all_gather(gathered, theta)
    identical = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    wl, single, nets1, _ = make_solver(key, 3001, data_parallel=False)
    single.fit(4, tqdm_file=None)
    theta1 = torch.cat([torch.as_tensor(p).reshape(-1) for p in get_params(nets1)]).cuda()
    fit_rel = float((theta - theta1).norm() / theta1.norm())
    losses_rel = float(np.max(np.abs(np.array(solver.metrics_history["train_loss"]) / np.array(single.metrics_history["train_loss"]) - 1.0)))
    dist.barrier()
    if rank == 0:
        with open(out, "w") as f:
            json.dump({"workload": key, "world": world, "points": n, "grad_rel": grad_rel, "sumsq_rel": sumsq_rel, "oneshot": oneshot,
                       "fit_theta_rel": fit_rel, "fit_loss_rel": losses_rel, "fit_ranks_identical": identical}, f)
        print("dp_nccl_worker", key, "oneshot", oneshot, "grad_rel", grad_rel, "sumsq_rel", sumsq_rel, "fit_theta_rel", fit_rel,
              "fit_loss_rel", losses_rel, "identical", identical, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)     # skip the NCCL teardown (can block after captured collectives)


if __name__ == "__main__":
    main()
