"""Optimizers that work on the engine's FLAT parameter / gradient buffers.

The fused engine keeps every ``nn.Parameter`` as a view into one flat ``theta`` tensor and every ``.grad`` as a view into one
flat ``grad`` tensor (``engine.FusedProblem._adopt_parameters``).  ``torch.optim.Adam`` over the individual parameters walks
a Python list of tensors and their state dicts every step; :class:`FlatAdam` applies the same update (Adam, Kingma & Ba;
the algorithm of ``torch.optim.Adam`` with ``amsgrad=False, weight_decay=0, maximize=False``, reference default
``solvers.py:182``) to the flat buffers with a handful of elementwise kernels, independent of the number of layers.

Opt-in::

    solver = Solver2D(...)
    solver.optimizer = FlatAdam.for_solver(solver, lr=1e-3)
"""
import math

import torch


class FlatAdam(torch.optim.Optimizer):
    """``capturable=True`` keeps the step counter, the learning rate and the bias corrections in device tensors, so that
    :meth:`step` is a fixed sequence of device operations (no host scalars that change from step to step) and can be
    recorded into the CUDA graph that already holds K0..K2b: a whole training step then replays as one graph."""

    def __init__(self, theta, grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        if theta.dim() != 1 or grad.shape != theta.shape:
            raise ValueError("FlatAdam expects the flat parameter buffer and its flat gradient buffer")
        if not 0.0 <= lr or not 0.0 <= eps or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        self._theta, self._grad = theta, grad
        holder = torch.nn.Parameter(theta.detach(), requires_grad=False)   # param_groups / state_dict plumbing
        super().__init__([holder], dict(lr=lr, betas=betas, eps=eps))
        self._m = torch.zeros_like(theta)
        self._v = torch.zeros_like(theta)
        self._t = 0
        self.capturable = capturable
        if capturable:
            # {step count, learning rate, launch ticket of pj_adam_step}: device doubles, so that a captured step replays
            self._state_dev = torch.zeros(3, dtype=torch.float64, device=theta.device)
            self._t_dev, self._lr_dev = self._state_dev[0], self._state_dev[1]
            self._lr_dev.fill_(float(lr))
            self._lr_host = float(lr)

    def sync_hyperparameters(self):
        """Capturable mode: push a learning rate edited through ``param_groups`` (schedulers, callbacks) to the device.
        Call outside the captured region, before a replay."""
        lr = float(self.param_groups[0]["lr"])
        if self.capturable and lr != self._lr_host:
            self._lr_dev.fill_(lr)
            self._lr_host = lr

    def _step_fused(self, loss=None, best_loss=None, best_theta=None):
        """The whole update as ONE launch of ``pj_adam_step`` (csrc/pinnjet_optim.cu); with ``best_theta`` the launch also
        keeps the best parameters seen so far (taken BEFORE the update, like the reference's ``_update_best``)."""
        import ctypes
        from .engine import load_library
        lib = load_library()
        if not hasattr(lib, "_adam_ready"):
            vp = ctypes.c_void_p
            lib.pj_adam_step.argtypes = [vp, vp, vp, vp, ctypes.c_int64, vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, vp, vp,
                                         vp, vp]
            lib.pj_adam_step.restype = ctypes.c_int
            lib._adam_ready = True
        b1, b2 = self.param_groups[0]["betas"]
        ptr = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        rc = lib.pj_adam_step(self._theta.data_ptr(), self._grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                              self._theta.numel(), self._state_dev.data_ptr(), b1, b2, self.param_groups[0]["eps"], ptr(loss),
                              ptr(best_loss), ptr(best_theta),
                              ctypes.c_void_p(torch.cuda.current_stream(self._theta.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"pj_adam_step failed ({rc})")

    def _step_on_device(self):
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        g = self._grad
        self._t_dev.add_(1.0)
        self._m.mul_(b1).add_(g, alpha=1.0 - b1)
        self._v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        bc1 = 1.0 - torch.pow(b1, self._t_dev)                     # 0-dim device tensors
        bc2 = 1.0 - torch.pow(b2, self._t_dev)
        denom = (self._v.sqrt() / bc2.sqrt().to(self._v.dtype)).add_(eps)
        self._theta.sub_((self._m / denom) * (self._lr_dev / bc1).to(self._theta.dtype))

    @classmethod
    def for_solver(cls, solver, **kw):
        return cls(solver.problem.theta, solver.problem.grad, **kw)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.capturable:
            self.sync_hyperparameters()
            self._step_on_device()
            self._t += 1
            return loss
        group = self.param_groups[0]
        lr, (b1, b2), eps = group["lr"], group["betas"], group["eps"]
        self._t += 1
        g = self._grad
        self._m.mul_(b1).add_(g, alpha=1.0 - b1)
        self._v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        bc1 = 1.0 - b1 ** self._t
        bc2 = 1.0 - b2 ** self._t
        denom = (self._v.sqrt() / math.sqrt(bc2)).add_(eps)
        self._theta.addcdiv_(self._m, denom, value=-lr / bc1)
        return loss

    def zero_grad(self, set_to_none=False):
        self._grad.zero_()

    # ---- checkpointing: the moments and the step count live outside ``self.state`` (flat buffers) ------------------------
    def state_dict(self):
        d = super().state_dict()
        d["flat_adam"] = {"step": int(self._t), "exp_avg": self._m.detach().clone(), "exp_avg_sq": self._v.detach().clone()}
        return d

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("flat_adam", None)
        super().load_state_dict(state_dict)
        if flat is not None:
            self._t = int(flat["step"])
            self._m.copy_(flat["exp_avg"].to(self._m.device, self._m.dtype))
            self._v.copy_(flat["exp_avg_sq"].to(self._v.device, self._v.dtype))
            if self.capturable:
                self._t_dev.fill_(float(self._t))
                self._lr_host = None                 # force the next sync to push the restored learning rate
                self.sync_hyperparameters()
