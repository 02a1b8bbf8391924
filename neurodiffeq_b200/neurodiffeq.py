"""``diff`` -- the reference's derivative API (neurodiffeq/neurodiffeq.py:6-82), jet-aware.

Inside a fused solver the arguments are :class:`~neurodiffeq_b200.symbolic.Sym` placeholders and ``diff`` is exact
symbolic differentiation that resolves to jet channels of the network output (no autograd graph, no per-order
backward sweep).  Called on ordinary tensors (user code outside a solver, e.g. post-processing) it keeps the
reference's eager semantics: ``order`` nested ``autograd.grad(create_graph=True)`` calls, unused input -> zeros.
"""
import torch

from . import symbolic as _sym
from ._compat import renamed_arguments


def _check_shapes(u, t):
    # same misuse -> same exception type and wording intent as reference neurodiffeq.py:52-59
    if len(u.shape) != 2 or len(t.shape) != 2 or u.shape[1] != 1 or t.shape[1] != 1:
        raise ValueError(f"Input shapes must both be (n_samples, 1) starting from neurodiffeq v0.2.0; \n"
                         f"got {tuple(u.shape)} (for dependent variable) and {tuple(t.shape)} (for independent "
                         f"variable). In most scenarios, consider reshaping inputs by `x = x.view(-1, 1)`")
    if u.shape != t.shape:
        raise ValueError(f"Input shapes must be the same shape starting from v0.2.0; "
                         f"got {tuple(u.shape)} != {tuple(t.shape)}")


@renamed_arguments(x="u")
def unsafe_diff(u, t, order=1):
    """Derivative without shape checks (reference neurodiffeq.py:6-34)."""
    if _sym.is_symbolic(u, t):
        return _sym.sym_diff(u, t, order=order)
    cur = u
    for _ in range(order):
        cur, = torch.autograd.grad(cur, t, grad_outputs=torch.ones_like(cur), create_graph=True, allow_unused=True)
        if cur is None:
            return torch.zeros_like(t, requires_grad=True)
        cur.requires_grad_()
    return cur


@renamed_arguments(x="u")
def safe_diff(u, t, order=1):
    """Derivative with the (n_samples, 1) shape contract (reference neurodiffeq.py:37-60)."""
    if not _sym.is_symbolic(u, t):
        _check_shapes(u, t)
    elif isinstance(u, torch.Tensor) or isinstance(t, torch.Tensor):
        raise ValueError("cannot mix traced symbols and eager tensors in diff()")
    return unsafe_diff(u, t, order=order)


@renamed_arguments(x="u")
def diff(u, t, order=1, shape_check=True):
    """d^order u / d t^order (reference neurodiffeq.py:63-82)."""
    return safe_diff(u, t, order=order) if shape_check else unsafe_diff(u, t, order=order)
