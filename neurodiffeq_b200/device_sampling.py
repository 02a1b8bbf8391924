"""Collocation point sampling on the device.  Opt-in: the host generators stay the default (BASELINE
north_star "generators.py point sampling stays on host").

A :class:`DeviceSampler` is built FROM a host generator object (``Generator1D``, ``Generator2D`` / ``Generator3D``,
``GeneratorSpherical``, ``StaticGenerator``, ``PredefinedGenerator`` and ``*`` ensembles of those) and draws from the same
law with the ``pj_sample`` kernel (``csrc/pinnjet_sample.cu``: Philox4x32-10, Box-Muller) straight into the static coordinate
buffers of the captured training step -- no host sampling, no pinned staging, no host-to-device copy per epoch.  The law is
a function of the GLOBAL row index, so under data parallelism a rank draws exactly its rows of the batch all ranks agree on.

The stream of random numbers differs from torch's CPU generator (statistically equivalent, not bit-equal); fixed-node
generators (``equally-spaced``, ``chebyshev*``, static / predefined points) are reproduced exactly.
Reference laws: generators.py:107-191 (1-D), :194-314 (grids: N(0, (step/4)^2) jitter), :572-655 (spherical).
"""
import ctypes

import torch

from . import generators as G
from .engine import PJ_MAX_COORDS, load_library

LAW_BASE, LAW_BASE_NORMAL, LAW_UNIFORM, LAW_SPHERICAL = 0, 1, 2, 3


class PjSampleLaw(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("coord", ctypes.c_int32), ("flag", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("p0", ctypes.c_float), ("p1", ctypes.c_float), ("div", ctypes.c_int64), ("mod", ctypes.c_int64),
                ("base", ctypes.c_void_p)]


class PjSampler(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_laws", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("law", PjSampleLaw * PJ_MAX_COORDS)]


def describe(gen):
    """Host generator -> list of law descriptions ``(kind, n_coords, payload)`` or ``None`` if the generator (or one of its
    parts) has no device law (``chebyshev2-noisy``, ``latin-hypercube`` redraws, filters, transforms, samplers ...)."""
    if isinstance(gen, G.SamplerGenerator):   # the solvers' wrapper: only reshapes the columns
        return describe(gen.generator)
    if isinstance(gen, G.EnsembleGenerator):
        out = []
        for g in gen.generators:
            part = describe(g)
            if part is None:
                return None
            out += part
        return out
    if isinstance(gen, G.MeshGenerator):   # 'ij' mesh of one-coordinate generators: a row's node on axis a is (row // div_a) % size_a
        parts = [describe(g) for g in gen.generators]
        if any(p is None or len(p) != 1 or p[0][1] != 1 for p in parts):
            return None
        out, div = [], 1
        for g, part in reversed(list(zip(gen.generators, parts))):
            kind, span, payload = part[0]
            out.append((kind, span, dict(payload, div=div, mod=int(g.size))))
            div *= int(g.size)
        return out[::-1]
    if isinstance(gen, (G.StaticGenerator, G.PredefinedGenerator)):
        return [(LAW_BASE, 1, dict(base=t.detach().reshape(-1).to(torch.float32))) for t in G._as_tuple(gen.get_examples())]
    if type(gen) is G.Generator1D:
        if gen.method == "uniform":
            return [(LAW_UNIFORM, 1, dict(p0=float(gen.t_min), p1=float(gen.t_max)))]
        if gen.method in ("equally-spaced", "log-spaced", "chebyshev", "chebyshev1", "chebyshev2"):
            return [(LAW_BASE, 1, dict(base=gen.getter().reshape(-1)))]
        if gen.method in ("equally-spaced-noisy", "log-spaced-noisy"):
            base = G.nodes_1d(gen.method[:-len("-noisy")], gen.t_min, gen.t_max, gen.size, None)()
            return [(LAW_BASE_NORMAL, 1, dict(base=base.reshape(-1), p0=float(gen.noise_std)))]
        return None
    if isinstance(gen, G._GridGenerator) and type(gen) in (G.Generator2D, G.Generator3D):
        if gen._static is None:
            return None
        if gen._std is None:
            return [(LAW_BASE, 1, dict(base=p.reshape(-1))) for p in gen._static]
        return [(LAW_BASE_NORMAL, 1, dict(base=p.reshape(-1), p0=float(s))) for p, s in zip(gen._static, gen._std)]
    if type(gen) is G.GeneratorSpherical:
        return [(LAW_SPHERICAL, 3, dict(p0=float(gen.r_min), p1=float(gen.r_max),
                                        flag=1 if gen.method == "equally-spaced-noisy" else 0))]
    return None


class DeviceSampler:
    """Draws the batch of ``gen`` on ``device``; ``sample_into(outs, first, n)`` fills rows ``[first, first + n)`` of the
    global batch into the float32 device tensors ``outs`` (one per coordinate, ``n`` elements each)."""

    def __init__(self, gen, device, seed=None):
        laws = describe(gen)
        if laws is None:
            raise ValueError(f"{gen!r} has no device sampling law")
        self.size = int(gen.size)
        self.n_coords = sum(k for _, k, _ in laws)
        if self.n_coords > PJ_MAX_COORDS:
            raise ValueError("too many coordinates")
        self.device = torch.device(device)
        self.lib = load_library()
        self.lib.pj_sample.argtypes = [ctypes.POINTER(PjSampler), ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.c_void_p, ctypes.c_void_p]
        self.lib.pj_sample.restype = ctypes.c_int
        self.spec = PjSampler()
        self.spec.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.spec.n_laws = len(laws)
        self._keep = []
        coord = 0
        for i, (kind, span, p) in enumerate(laws):
            law = self.spec.law[i]
            law.kind, law.coord, law.flag = kind, coord, int(p.get("flag", 0))
            law.p0, law.p1 = float(p.get("p0", 0.0)), float(p.get("p1", 0.0))
            law.div, law.mod = int(p.get("div", 1)), int(p.get("mod", 0))
            if "base" in p:
                if p["base"].numel() != (law.mod if law.mod > 0 else self.size):
                    raise ValueError("base nodes do not match the generator size")
                t = p["base"].to(self.device, torch.float32).contiguous()
                self._keep.append(t)
                law.base = t.data_ptr()
            coord += span
        self.state = torch.zeros(2, dtype=torch.int64, device=self.device)     # {call number, launch ticket}

    @classmethod
    def from_generator(cls, gen, device, seed=None):
        """``DeviceSampler`` of ``gen`` or ``None`` when the generator has no device law."""
        return cls(gen, device, seed) if describe(gen) is not None else None

    def sample_into(self, outs, first=0, n=None):
        n = self.size - first if n is None else n
        if len(outs) != self.n_coords or any(o.numel() < n or o.dtype != torch.float32 or not o.is_cuda for o in outs):
            raise ValueError("sample_into: one float32 device tensor of >= n elements per coordinate")
        ptrs = (ctypes.c_void_p * PJ_MAX_COORDS)(*([o.data_ptr() for o in outs] + [None] * (PJ_MAX_COORDS - len(outs))))
        rc = self.lib.pj_sample(ctypes.byref(self.spec), int(first), int(n), ptrs, self.state.data_ptr(),
                                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"pj_sample failed ({rc})")
        return outs

    def get_examples(self):
        """Generator protocol (whole batch, fresh device tensors) -- for tests and eager use."""
        outs = [torch.empty(self.size, dtype=torch.float32, device=self.device) for _ in range(self.n_coords)]
        self.sample_into(outs, 0, self.size)
        return outs[0] if len(outs) == 1 else tuple(outs)
