"""Network definitions accepted by the fused path (reference neurodiffeq/networks.py:6-70, 142-152).

``FCNN`` keeps the reference's module layout -- an ``nn.Sequential`` called ``NN`` of ``Linear, actv, ..., Linear`` with
parameters at ``NN.{0,2,4,...}.{weight,bias}`` -- so state dicts, optimizers, checkpoints and ``deepcopy`` of user code
keep working; the CUDA engine reads the weights of exactly this structure (any ``nn.Sequential`` alternating
``nn.Linear`` with ``nn.Tanh`` / ``SinActv`` is accepted, see ``engine.describe_network``).
"""
from warnings import warn

import torch
import torch.nn as nn


class SinActv(nn.Module):
    """sin activation (reference networks.py:142-152)."""

    def forward(self, input_):
        return torch.sin(input_)


class FCNN(nn.Module):
    """Fully connected network; defaults (32, 32) hidden units and tanh like the reference (networks.py:52-53)."""

    def __init__(self, n_input_units=1, n_output_units=1, n_hidden_units=None, n_hidden_layers=None,
                 actv=nn.Tanh, hidden_units=None):
        super().__init__()
        if n_hidden_units is not None or n_hidden_layers is not None:  # deprecated pair, reference networks.py:32-50
            n_hidden_units = 32 if n_hidden_units is None else n_hidden_units
            n_hidden_layers = 1 if n_hidden_layers is None else n_hidden_layers
            if hidden_units is None:
                hidden_units = tuple(n_hidden_units for _ in range(n_hidden_layers + 1))
                warn(f"`n_hidden_units` and `n_hidden_layers` are deprecated, pass `hidden_units={hidden_units}`",
                     FutureWarning)
            else:
                warn(f"Ignoring `n_hidden_units` and `n_hidden_layers` in favor of `hidden_units={hidden_units}`",
                     FutureWarning)
        if hidden_units is None:
            hidden_units = (32, 32)
        hidden_units = tuple(hidden_units)
        widths = (n_input_units,) + hidden_units
        layers = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers.append(nn.Linear(fan_in, fan_out))
            layers.append(actv())
        layers.append(nn.Linear(widths[-1], n_output_units))
        self.NN = nn.Sequential(*layers)

    def forward(self, t):
        return self.NN(t)


# ----------------------------------------------------------------------------------------------------------------------
# The remaining modules of the reference's networks.py.  They are ordinary torch modules (usable in eager code, same
# semantics as the reference); the fused engine has no jet rule for them yet and says so when a solver is built with one.
# ----------------------------------------------------------------------------------------------------------------------
class Resnet(nn.Module):
    """``skip_connection(t) + residual(t)``: a bias-free Linear from input to output beside an FCNN (networks.py:73-106)."""

    def __init__(self, n_input_units=1, n_output_units=1, n_hidden_units=None, n_hidden_layers=None, actv=nn.Tanh,
                 hidden_units=(32, 32)):
        super().__init__()
        self.residual = FCNN(n_input_units=n_input_units, n_output_units=n_output_units, n_hidden_units=n_hidden_units,
                             n_hidden_layers=n_hidden_layers, actv=actv, hidden_units=hidden_units)
        self.skip_connection = nn.Linear(n_input_units, n_output_units, bias=False)

    def forward(self, t):
        return self.skip_connection(t) + self.residual(t)


class MonomialNN(nn.Module):
    """Feature map ``x -> [x**d for d in degrees]`` concatenated along dim 1 (networks.py:109-139)."""

    def __init__(self, degrees):
        super().__init__()
        if isinstance(degrees, int):
            degrees = list(range(1, degrees + 1))
        self.degrees = tuple(degrees)
        if len(self.degrees) == 0:
            raise ValueError("No degrees used, check `degrees` argument again")
        if 0 in self.degrees:
            warn("One of the degrees is 0 which might introduce redundant features")
        if len(set(self.degrees)) < len(self.degrees):
            warn(f"Duplicate degrees found: {self.degrees}")

    def forward(self, x):
        return torch.cat([x ** d for d in self.degrees], dim=1)

    def __repr__(self):
        return f"{self.__class__.__name__}(degrees={self.degrees})"

    __str__ = __repr__


class Swish(nn.Module):
    """``x * sigmoid(beta * x)`` with an optionally trainable ``beta`` (networks.py:155-174)."""

    def __init__(self, beta=1.0, trainable=False):
        super().__init__()
        self.trainable = trainable
        self.beta = nn.Parameter(torch.tensor(float(beta))) if trainable else float(beta)

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class APTx(nn.Module):
    """``(alpha + tanh(beta * x)) * gamma * x`` with optionally trainable scalars (networks.py:177-208)."""

    def __init__(self, alpha=1.0, beta=1.0, gamma=0.5, trainable=False):
        super().__init__()
        self.trainable = trainable
        wrap = (lambda v: nn.Parameter(torch.tensor(float(v)))) if trainable else float
        self.alpha, self.beta, self.gamma = wrap(alpha), wrap(beta), wrap(gamma)

    def forward(self, x):
        return (self.alpha + torch.tanh(self.beta * x)) * self.gamma * x
