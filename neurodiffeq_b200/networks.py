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
