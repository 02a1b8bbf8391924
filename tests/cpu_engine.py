"""TEST INFRASTRUCTURE: a CPU stand-in for ``neurodiffeq_b200.engine.FusedProblem`` built on the numpy mirror of the
kernels (oracle/jet_numpy.py, float64).  It exists so that the HOST logic of the solvers -- epoch loop, loss bookkeeping,
custom / named losses, best-network tracking, solutions, residuals -- is exercised by the CPU test suite; the product
itself never imports it (the real engine fails loudly without CUDA)."""
import numpy as np
import torch

from neurodiffeq_b200.engine import pad_scheme, combine_seconds
from neurodiffeq_b200.tracing import TracedProblem
from oracle import jet_numpy


class CpuFusedProblem:
    def __init__(self, nets, conditions, diff_eqs, n_coords, coords_for_condition=None, device=None, aux_outputs=None,
                 enforce=None):
        self.device = torch.device("cpu")
        self.tp = TracedProblem(nets, conditions, diff_eqs, n_coords, coords_for_condition, pad_scheme=pad_scheme,
                                combine_seconds=combine_seconds, aux_outputs=aux_outputs, enforce=enforce)
        self.n_coords, self.n_funcs, self.n_eq = n_coords, self.tp.n_funcs, self.tp.n_eq
        params, seen = [], set()
        for nd in self.tp.nets:
            nd.module.to(dtype=torch.float64)
            for p in nd.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        n_theta = sum(p.numel() for p in params)
        self.theta = torch.empty(n_theta, dtype=torch.float64)
        self.gradbuf = torch.zeros(n_theta + 1, dtype=torch.float64)
        self.grad, self.sumsq = self.gradbuf[:n_theta], self.gradbuf[n_theta:]
        self.params, self.offsets, off = params, [], 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.theta[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.theta[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                self.offsets.append(off)
                off += n
        self.n_theta = n_theta
        self.kernel_launches = 0

    def parameters_linked(self):
        return True

    def relink(self):
        pass

    jit_reason = "CPU stand-in"

    def enable_jit(self, strict=False):
        return False

    def pack(self, zero_gradbuf=False):
        if zero_gradbuf:
            self.gradbuf.zero_()

    def _per_instance(self):
        by_param = {id(p): p.detach().numpy() for p in self.params}
        return [[by_param[id(q)] for q in nd.parameters()] for nd in self.tp.nets]

    @staticmethod
    def _np(coords):
        return np.stack([c.detach().cpu().double().reshape(-1).numpy() for c in coords])

    def forward(self, coords, want_u=True, want_residual=True, want_sumsq=False, repack=True):
        out = jet_numpy.run_traced(self.tp, self._per_instance(), self._np(coords), want_grad=False)
        if want_sumsq:
            self.sumsq.zero_()
            self.sumsq += float((out["residual"] ** 2).sum())
        return (torch.from_numpy(out["u"]) if want_u else None, torch.from_numpy(out["residual"]) if want_residual else None,
                self.sumsq if want_sumsq else None)

    def residual_grad(self, coords, n_global=None, want_residual=False, rbar=None, sumsq_out=None, repack=True, ubar=None):
        out = jet_numpy.run_traced(self.tp, self._per_instance(), self._np(coords), n_global=n_global,
                                   rbar=None if rbar is None else rbar.detach().numpy(),
                                   ubar=None if ubar is None else ubar.detach().numpy())
        with torch.no_grad():
            for g, off in zip(out["grads"], self.offsets):   # accumulate, like loss.backward()
                self.grad[off:off + g.size] += torch.from_numpy(np.ascontiguousarray(g).reshape(-1))
            if sumsq_out is None:
                sumsq_out = self.sumsq
                sumsq_out.zero_()
            sumsq_out += float((out["residual"] ** 2).sum())
        return sumsq_out, (torch.from_numpy(out["residual"]) if want_residual else None)

    def grads_as_list(self):
        return [self.grad[o:o + p.numel()].view(p.shape).detach().numpy().copy() for p, o in zip(self.params, self.offsets)]

    def residual_grad_graphed(self, coords, n_global=None, train=True, zero_gradbuf=False):
        if zero_gradbuf and train:
            self.gradbuf.zero_()
        if train:
            self.residual_grad(coords, n_global=n_global, sumsq_out=self.sumsq)
        else:
            r = jet_numpy.run_traced(self.tp, self._per_instance(), self._np(coords), want_grad=False)["residual"]
            with torch.no_grad():
                self.sumsq += float((r ** 2).sum())
        return self.sumsq
