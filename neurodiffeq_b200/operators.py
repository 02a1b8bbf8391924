"""Vector-calculus operators of the reference (neurodiffeq/operators.py:15-432), written once over ``diff`` so that
they work on traced symbols (fused path: they expand to jet-channel algebra) and on eager
tensors alike.  Formulas follow the textbook definitions the reference's tests pin
(tests/test_operators_cartesian.py:51-111, test_operators_spherical.py:78-147, test_operators_cylindrical.py:64-110).
"""
import torch

from .neurodiffeq import safe_diff as diff
from . import symbolic as _sym


def _sin(x):
    return x.sin()


def _cos(x):
    return x.cos()


def _split_u_x(*us_xs):
    n = len(us_xs)
    if n == 0 or n % 2 != 0:
        raise RuntimeError("Number of us and xs must be equal and positive")  # reference operators.py:7-12
    return us_xs[:n // 2], us_xs[n // 2:]


def grad(u, *xs):
    """[du/dx_i]; a coordinate ``u`` does not depend on gives zeros (operators.py:15-33)."""
    if _sym.is_symbolic(u, *xs):
        return [diff(u, x) for x in xs]
    gs = torch.autograd.grad(u, xs, grad_outputs=torch.ones_like(u), create_graph=True, allow_unused=True)
    return [torch.zeros_like(x, requires_grad=True) if g is None else g.requires_grad_(True) for x, g in zip(xs, gs)]


def div(*us_xs):
    us, xs = _split_u_x(*us_xs)
    return sum(diff(u, x) for u, x in zip(us, xs))


def curl(u_x, u_y, u_z, x, y, z):
    dxy, dxz = grad(u_x, y, z)
    dyx, dyz = grad(u_y, x, z)
    dzx, dzy = grad(u_z, x, y)
    return dzy - dyz, dxz - dzx, dyx - dxy


def laplacian(u, *xs):
    return sum(diff(g, x) for g, x in zip(grad(u, *xs), xs))


def vector_laplacian(u_x, u_y, u_z, x, y, z):
    return laplacian(u_x, x, y, z), laplacian(u_y, x, y, z), laplacian(u_z, x, y, z)


# ---- spherical (r, theta = co-latitude, phi = longitude) ------------------------------------------------------------
def spherical_curl(u_r, u_theta, u_phi, r, theta, phi):
    ur_th, ur_ph = grad(u_r, theta, phi)
    uth_r, uth_ph = grad(u_theta, r, phi)
    uph_r, uph_th = grad(u_phi, r, theta)
    s, c = _sin(theta), _cos(theta)
    return ((uph_th + (u_phi * c - uth_ph) / s) / r,
            (ur_ph / s - u_phi) / r - uph_r,
            uth_r + (u_theta - ur_th) / r)


def spherical_grad(u, r, theta, phi):
    u_r, u_th, u_ph = grad(u, r, theta, phi)
    return u_r, u_th / r, u_ph / (r * _sin(theta))


def spherical_div(u_r, u_theta, u_phi, r, theta, phi):
    s = _sin(theta)
    return (diff(u_r * r ** 2, r) / r + (diff(u_theta * s, theta) + diff(u_phi, phi)) / s) / r


def spherical_laplacian(u, r, theta, phi):
    u_r, u_th, u_ph = grad(u, r, theta, phi)
    s = _sin(theta)
    r2 = r ** 2
    return (diff(r2 * u_r, r) + diff(s * u_th, theta) / s + diff(u_ph, phi) / s ** 2) / r2


def spherical_vector_laplacian(u_r, u_theta, u_phi, r, theta, phi):
    ur_th, ur_ph = grad(u_r, theta, phi)
    uth_th, uth_ph = grad(u_theta, theta, phi)
    uph_ph = diff(u_phi, phi)
    s, c = _sin(theta), _cos(theta)
    r2 = r ** 2
    lap = lambda f: spherical_laplacian(f, r, theta, phi)  # noqa: E731
    return (lap(u_r) - 2 * (u_r + uth_th + (c * u_theta + uph_ph) / s) / r2,
            lap(u_theta) + (2 * ur_th - (u_theta + 2 * c * uph_ph) / s ** 2) / r2,
            lap(u_phi) + ((2 * c * uth_ph - u_phi) / s + 2 * ur_ph) / (s * r2))


def spherical_to_cartesian(r, theta, phi):
    s = _sin(theta)
    return r * s * _cos(phi), r * s * _sin(phi), r * _cos(theta)


def cartesian_to_spherical(x, y, z):
    return (torch.sqrt(x ** 2 + y ** 2 + z ** 2), torch.atan2(torch.sqrt(x ** 2 + y ** 2), z), torch.atan2(y, x))


# ---- cylindrical (rho, phi, z) ---------------------------------------------------------------------------------------
def cylindrical_grad(u, rho, phi, z):
    u_rho, u_phi, u_z = grad(u, rho, phi, z)
    return u_rho, u_phi / rho, u_z


def cylindrical_div(u_rho, u_phi, u_z, rho, phi, z):
    return diff(u_rho, rho) + (u_rho + diff(u_phi, phi)) / rho + diff(u_z, z)


def cylindrical_curl(u_rho, u_phi, u_z, rho, phi, z):
    urho_phi, urho_z = grad(u_rho, phi, z)
    uphi_rho, uphi_z = grad(u_phi, rho, z)
    uz_rho, uz_phi = grad(u_z, rho, phi)
    return uz_phi / rho - uphi_z, urho_z - uz_rho, uphi_rho + (u_phi - urho_phi) / rho


def cylindrical_laplacian(u, rho, phi, z):
    u_rho, u_phi, u_z = grad(u, rho, phi, z)
    return diff(u_rho, rho) + u_rho / rho + diff(u_phi, phi) / rho ** 2 + diff(u_z, z)


def cylindrical_vector_laplacian(u_rho, u_phi, u_z, rho, phi, z):
    lap = lambda f: cylindrical_laplacian(f, rho, phi, z)  # noqa: E731
    return (lap(u_rho) - (u_rho + 2 * diff(u_phi, phi)) / rho ** 2,
            lap(u_phi) + (2 * diff(u_rho, phi) - u_phi) / rho ** 2,
            lap(u_z))


def cylindrical_to_cartesian(rho, phi, z):
    return rho * _cos(phi), rho * _sin(phi), z


def cartesian_to_cylindrical(x, y, z):
    return torch.sqrt(x ** 2 + y ** 2), torch.atan2(y, x), z
