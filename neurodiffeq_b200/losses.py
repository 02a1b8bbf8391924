"""Loss functions of the reference (neurodiffeq/losses.py:5-35) by name.

``l2`` is the fused mean-squared residual.  ``l1`` / ``infinity`` go through the solver's custom-loss path (autograd gives
dL/dr on the small residual matrix, the kernels do the rest).  ``h1`` adds, per coordinate, the derivative of the
(equation-summed) residual -- ``grad(residual, *coords)`` of losses.py:18 differentiates the SUM of the residual columns
-- and averages the squares of all ``n_eq + d`` columns: the fused solvers obtain those columns by differentiating the
traced residuals symbolically, i.e. the same fused mean-squared path over an augmented residual list.  ``h1 semi`` uses the
derivative columns only; the user's residuals stay available (``get_residuals``) as auxiliary outputs of the forward kernel.
"""
import torch


def _l1_norm(residual, funcs, coords):
    return torch.abs(residual).mean()


def _l2_norm(residual, funcs, coords):
    return (residual ** 2).mean()


def _infinity_norm(residual, funcs, coords):
    return residual.abs().max(dim=1)[0].mean()


def _h1_norm(residual, funcs, coords):
    from .operators import grad
    g = grad(residual, *coords)
    return (torch.cat([residual, *g], dim=1) ** 2).mean()


def _h1_semi_norm(residual, funcs, coords):
    from .operators import grad
    return (torch.cat(grad(residual, *coords), dim=1) ** 2).mean()


_losses = {"l1": _l1_norm, "l2": _l2_norm, "infinity": _infinity_norm, "h1": _h1_norm, "h1 semi": _h1_semi_norm}


def h1_semi_rows(diff_eqs, n_funcs):
    """rows of the 'h1 semi' loss: d(sum of the equations)/d(coord) for every coordinate, without the equations"""
    from .neurodiffeq import diff

    def rows(*variables):
        res = diff_eqs(*variables)
        res = list(res) if isinstance(res, (list, tuple)) else [res]
        total = res[0]
        for r in res[1:]:
            total = total + r
        return [diff(total, c) for c in variables[n_funcs:]]
    return rows


def h1_rows(diff_eqs, n_funcs):
    """``diff_eqs`` -> callable returning the residual rows whose mean square IS the 'h1' loss: the equations, then
    d(sum of the equations)/d(coord) for every coordinate (used by the fused solvers on traced symbols)."""
    from .neurodiffeq import diff

    def rows(*variables):
        res = diff_eqs(*variables)
        res = list(res) if isinstance(res, (list, tuple)) else [res]
        total = res[0]
        for r in res[1:]:
            total = total + r
        return res + [diff(total, c) for c in variables[n_funcs:]]
    return rows
