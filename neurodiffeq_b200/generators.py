"""Host-side collocation point sampling (reference neurodiffeq/generators.py) -- "generators.py point sampling stays on
host" (BASELINE.json north_star).

The fused solvers only need ``generator.get_examples()`` (tensors of ``size`` points per coordinate) and
``generator.size`` (reference solvers.py:49-52), so any reference generator object works as well.  This module re-states
the commonly used ones so that user code runs with an import-root change only: 1-D / 2-D / 3-D / spherical samplers
the N-D tensor-product sampler, the ``+`` (concat), ``*`` (ensemble) and ``^`` (mesh) combinators and the
wrappers (static, predefined, transform, filter, resample, batch).  Samples are float32 CPU tensors WITHOUT
``requires_grad``: the fused path never builds an autograd graph over coordinates.

Sampling laws follow the reference: noisy grids add N(0, (step/4)^2) noise (generators.py:149-158, 253-266), spherical
points draw r^2 uniformly and directions from normalised random octant vectors (:603-646).
"""
import math

import numpy as np
import torch

_F = torch.float32


def _cheb1(a, b, n):
    x = torch.cos((torch.arange(n, dtype=_F) + 0.5) / n * math.pi)
    return ((a + b) + (b - a) * x) / 2


def _cheb2(a, b, n, jitter=False):
    k = torch.arange(n, dtype=_F)
    if jitter:
        k = k + (torch.rand(n) * 2 - 1)
    return ((a + b) + (b - a) * torch.cos(k / float(n - 1) * math.pi)) / 2


def _lhs(a, b, n):
    edges = torch.linspace(a, b, n + 1, dtype=_F)
    pts = edges[:-1] + torch.rand(n) * (edges[1] - edges[0])
    return pts[torch.randperm(n)]


def _log_bounds(lo, hi, who):
    if lo <= 0 or hi <= 0:
        raise ValueError(f"the interval [{lo}, {hi}] cannot be used for log-sampling in {who}; "
                         f"pass positive bounds (did you mean [{10 ** lo}, {10 ** hi}]?)")
    return math.log10(lo), math.log10(hi)


def nodes_1d(method, lo, hi, n, noise_std, who="Generator1D"):
    """Returns a zero-argument sampler for one coordinate."""
    if method == "uniform":
        return lambda: torch.rand(n) * (hi - lo) + lo
    if method in ("equally-spaced", "equally-spaced-noisy"):
        base = torch.linspace(lo, hi, n, dtype=_F)
    elif method in ("log-spaced", "log-spaced-noisy"):
        base = torch.logspace(*_log_bounds(lo, hi, who), n, dtype=_F)
    elif method in ("chebyshev", "chebyshev1"):
        base = _cheb1(lo, hi, n)
    elif method == "chebyshev2":
        base = _cheb2(lo, hi, n)
    elif method == "chebyshev2-noisy":
        return lambda: _cheb2(lo, hi, n, jitter=True)
    elif method == "latin-hypercube":
        return lambda: _lhs(lo, hi, n)
    else:
        raise ValueError(f"Unknown method: {method}")
    if method.endswith("-noisy"):
        return lambda: torch.normal(mean=base, std=noise_std)
    return lambda: base


class BaseGenerator:
    """``get_examples()`` + ``size``; ``+`` concatenates, ``*`` ensembles, ``^`` meshes (generators.py:46-66)."""

    def __init__(self):
        self.size = None

    def get_examples(self):
        raise NotImplementedError

    @staticmethod
    def check_generator(obj):
        if not isinstance(obj, BaseGenerator):
            raise ValueError(f"{obj} is not a generator")

    def __add__(self, other):
        self.check_generator(other)
        return ConcatGenerator(self, other)

    def __mul__(self, other):
        self.check_generator(other)
        return EnsembleGenerator(self, other)

    def __xor__(self, other):
        self.check_generator(other)
        return MeshGenerator(self, other)

    def __repr__(self):
        return f"{self.__class__.__name__}(size={self.size})"


class Generator1D(BaseGenerator):
    def __init__(self, size, t_min=0.0, t_max=1.0, method="uniform", noise_std=None):
        super().__init__()
        self.size, self.t_min, self.t_max, self.method = size, t_min, t_max, method
        self.noise_std = noise_std if noise_std else ((t_max - t_min) / size) / 4.0
        self.getter = nodes_1d(method, t_min, t_max, size, self.noise_std, self.__class__.__name__)

    def get_examples(self):
        return self.getter()


def _grid_axes(method, lo, hi, grid, noise_std, who):
    """Per-axis base nodes for the tensor-product generators; noise (if any) is added to the flattened mesh."""
    noisy = method == "equally-spaced-noisy"
    axis_method = "equally-spaced" if noisy else method
    samplers = [nodes_1d(axis_method, lo[d], hi[d], grid[d], None, who) for d in range(len(grid))]
    std = None
    if noisy:
        std = tuple(noise_std) if noise_std else tuple(((hi[d] - lo[d]) / grid[d]) / 4.0 for d in range(len(grid)))
    return samplers, std


class _GridGenerator(BaseGenerator):
    """Tensor-product points ('ij' mesh of per-axis nodes), optionally jittered: Generator2D / Generator3D."""

    def __init__(self, grid, lo, hi, method, noise_std):
        super().__init__()
        if method not in ("equally-spaced", "equally-spaced-noisy", "chebyshev", "chebyshev1", "chebyshev2",
                          "chebyshev2-noisy", "latin-hypercube"):
            raise ValueError(f"Unknown method: {method}")
        self.grid, self.method = tuple(grid), method
        self.size = int(np.prod(self.grid))
        self._samplers, self._std = _grid_axes(method, lo, hi, self.grid, noise_std, self.__class__.__name__)
        self._static = None
        if method in ("equally-spaced", "equally-spaced-noisy", "chebyshev", "chebyshev1", "chebyshev2",
                      "latin-hypercube"):  # axes drawn once (the reference also fixes the LHS draw at construction)
            self._static = self._mesh()

    def _mesh(self):
        axes = [s() for s in self._samplers]
        return tuple(m.flatten() for m in torch.meshgrid(*axes, indexing="ij"))

    def get_examples(self):
        pts = self._static if self._static is not None else self._mesh()
        if self._std is not None:
            pts = tuple(torch.normal(mean=p, std=s) for p, s in zip(pts, self._std))
        return pts


class Generator2D(_GridGenerator):
    def __init__(self, grid=(10, 10), xy_min=(0.0, 0.0), xy_max=(1.0, 1.0), method="equally-spaced-noisy",
                 xy_noise_std=None):
        super().__init__(grid, xy_min, xy_max, method, xy_noise_std)
        self.xy_min, self.xy_max, self.xy_noise_std = xy_min, xy_max, xy_noise_std


class Generator3D(_GridGenerator):
    def __init__(self, grid=(10, 10, 10), xyz_min=(0.0, 0.0, 0.0), xyz_max=(1.0, 1.0, 1.0),
                 method="equally-spaced-noisy"):
        super().__init__(grid, xyz_min, xyz_max, method, None)
        self.xyz_min, self.xyz_max = xyz_min, xyz_max


class GeneratorSpherical(BaseGenerator):
    """(r, theta, phi) with theta the co-latitude; directions are never exactly on the poles (generators.py:622-646)."""

    def __init__(self, size, r_min=0., r_max=1., method="equally-spaced-noisy"):
        super().__init__()
        if r_min < 0 or r_max < r_min:
            raise ValueError(f"Illegal range [{r_min}, {r_max}]")
        if method not in ("equally-spaced-noisy", "equally-radius-noisy"):
            raise ValueError(f"Unknown method: {method}")
        self.size, self.r_min, self.r_max, self.method = size, r_min, r_max, method

    def _radius(self):
        u = torch.rand(self.size)
        if self.method == "equally-spaced-noisy":  # r^2 uniform
            return torch.sqrt((self.r_max ** 2 - self.r_min ** 2) * u + self.r_min ** 2)
        return (self.r_max - self.r_min) * u + self.r_min

    def get_examples(self):
        w = torch.rand(3, self.size)
        v = torch.sqrt(w / w.sum(dim=0, keepdim=True)) + 1e-6          # point of the positive octant
        v = v * (torch.randint(0, 2, (3, self.size), dtype=v.dtype) * 2 - 1)   # random octant
        theta = torch.acos(v[2])
        phi = math.pi - torch.atan2(v[1], v[0])                        # [0, 2 pi)
        return self._radius(), theta, phi


def _as_tuple(ex):
    if isinstance(ex, torch.Tensor):
        return (ex,)
    return tuple(ex)


class ConcatGenerator(BaseGenerator):
    def __init__(self, *generators):
        super().__init__()
        self.generators = generators
        self.size = sum(g.size for g in generators)

    def get_examples(self):
        parts = [g.get_examples() for g in self.generators]
        if isinstance(parts[0], torch.Tensor):
            return torch.cat(parts)
        return [torch.cat(seg) for seg in zip(*parts)]


class EnsembleGenerator(BaseGenerator):
    def __init__(self, *generators):
        super().__init__()
        self.size = generators[0].size
        for i, g in enumerate(generators):
            if g.size != self.size:
                raise ValueError(f"gens[{i}].size ({g.size}) != gens[0].size ({self.size})")
        self.generators = generators

    def get_examples(self):
        out = tuple(t for g in self.generators for t in _as_tuple(g.get_examples()))
        return out[0] if len(out) == 1 else out


class MeshGenerator(BaseGenerator):
    def __init__(self, *generators):
        super().__init__()
        self.generators = []
        for g in generators:
            self.generators += list(g.generators) if isinstance(g, MeshGenerator) else [g]
        self.size = int(np.prod([g.size for g in self.generators]))

    def get_examples(self):
        axes = tuple(t for g in self.generators for t in _as_tuple(g.get_examples()))
        if len(axes) == 1:
            return axes[0]
        return tuple(m.flatten() for m in torch.meshgrid(*axes, indexing="ij"))


class StaticGenerator(BaseGenerator):
    """Samples once at construction, returns the same points forever (generators.py:691-714)."""

    def __init__(self, generator):
        super().__init__()
        self.size = generator.size
        self.examples = generator.get_examples()

    def get_examples(self):
        return self.examples


class PredefinedGenerator(BaseGenerator):
    """Returns user supplied points (generators.py:717-755)."""

    def __init__(self, *xs):
        super().__init__()
        self.size = len(xs[0])
        for x in xs:
            if len(x) != self.size:
                raise ValueError("tensors of different lengths encountered")
        self.xs = [x if isinstance(x, torch.Tensor) else torch.tensor(x) for x in xs]
        self.xs = [x.detach().to(_F).reshape(-1) for x in self.xs]

    def get_examples(self):
        return self.xs[0] if len(self.xs) == 1 else tuple(self.xs)


class TransformGenerator(BaseGenerator):
    """Applies per-coordinate maps (``transforms``: a list, ``None`` entries = identity) or one joint map (``transform``:
    ``f(*xs) -> tuple``) to another generator's samples (generators.py:758-801)."""

    def __init__(self, generator, transforms=None, transform=None):
        super().__init__()
        if transforms is not None and transform is not None:
            raise ValueError("transform and transforms cannot be both specified")
        self.generator, self.size = generator, generator.size
        self._per_axis = None if transforms is None else [t if t is not None else (lambda x: x) for t in transforms]
        self._joint = transform

    def get_examples(self):
        xs = self.generator.get_examples()
        single = isinstance(xs, torch.Tensor)
        if self._per_axis is not None:
            return self._per_axis[0](xs) if single else tuple(t(x) for t, x in zip(self._per_axis, xs))
        if self._joint is not None:
            return self._joint(xs) if single else self._joint(*xs)
        return xs


class FilterGenerator(BaseGenerator):
    """Keeps the samples where ``filter_fn(list_of_tensors)`` (a boolean mask) is true (generators.py:904-953);
    ``size`` follows the number of survivors unless ``update_size=False``."""

    def __init__(self, generator, filter_fn, size=None, update_size=True):
        super().__init__()
        self.generator, self.filter_fn, self.update_size = generator, filter_fn, update_size
        self.size = generator.size if size is None else size

    def get_examples(self):
        xs = list(_as_tuple(self.generator.get_examples()))
        keep = self.filter_fn(xs)
        xs = [x[keep] for x in xs]
        if self.update_size:
            self.size = len(xs[0])
        return xs[0] if len(xs) == 1 else xs


class ResampleGenerator(BaseGenerator):
    """A random subset (``replacement=False``: a permutation prefix) or bootstrap sample (``True``) of ``size`` points
    of another generator's batch; the same rows are taken from every coordinate (generators.py:956-993)."""

    def __init__(self, generator, size=None, replacement=False):
        super().__init__()
        self.generator, self.replacement = generator, replacement
        self.size = generator.size if size is None else size

    def get_examples(self):
        m = self.generator.size
        rows = torch.randint(m, (self.size,)) if self.replacement else torch.randperm(m)[:self.size]
        xs = self.generator.get_examples()
        return xs[rows] if isinstance(xs, torch.Tensor) else [x[rows] for x in xs]


class BatchGenerator(BaseGenerator):
    """Serves another generator's samples ``batch_size`` at a time from a cache that is refilled when it runs short
    (generators.py:996-1043)."""

    def __init__(self, generator, batch_size):
        super().__init__()
        if generator.size <= 0:
            raise ValueError(f"generator has size {generator.size} <= 0")
        self.generator, self.size = generator, batch_size
        self._cache = list(_as_tuple(generator.get_examples()))

    def get_examples(self):
        while len(self._cache[0]) < self.size:
            more = _as_tuple(self.generator.get_examples())
            self._cache = [torch.cat([old, new]) for old, new in zip(self._cache, more)]
        batch = [x[:self.size] for x in self._cache]
        self._cache = [x[self.size:] for x in self._cache]
        return batch[0] if len(batch) == 1 else batch


class GeneratorND(BaseGenerator):
    """Tensor-product points in N dimensions with a sampling law per axis (generators.py:419-570): 'equally-spaced',
    'uniform' (drawn once), 'log-spaced', 'exp-spaced' (equally spaced in ``base**x``), 'chebyshev'/'chebyshev1',
    'chebyshev2'; optional ``cut=(lo, hi)`` slices per axis; with ``noisy`` every call adds N(0, std^2) noise to the mesh
    (std per axis: ``r_noise_std`` or a quarter grid step, scaled by the node for the log / exp laws, 0 for 'uniform');
    ``abs_value`` folds the noisy points to non-negative values."""

    def __init__(self, grid=(10, 10), r_min=(0.0, 0.0), r_max=(1.0, 1.0), methods=("equally-spaced", "equally-spaced"),
                 noisy=True, r_noise_std=None, **kwargs):
        super().__init__()
        self.grid, self.r_min, self.r_max, self.methods = grid, r_min, r_max, methods
        self.noisy, self.r_noise_std = noisy, r_noise_std
        seq = lambda v: (v,) if isinstance(v, (int, float, str)) else tuple(v)  # noqa: E731
        grid, lo, hi, methods = seq(grid), seq(r_min), seq(r_max), seq(methods)
        stds = None if not r_noise_std else seq(r_noise_std)
        n_axes = len(grid)
        self.size = int(np.prod(grid))
        cut = kwargs.pop("cut", tuple((None, None) for _ in range(n_axes)))
        base = kwargs.pop("base", tuple(10 for _ in range(n_axes)))
        abs_value = kwargs.pop("abs_value", False)
        if kwargs:
            raise ValueError(f"Unknown keyword argument(s): {list(kwargs.keys())}")
        base = seq(base)
        if cut[0] is None or isinstance(cut[0], (int, float)):
            cut = (cut,)
        nodes, sigmas = [], []
        for d in range(n_axes):
            a, b, n, law = lo[d], hi[d], grid[d], methods[d]
            sd = stds[d] if stds else ((b - a) / n) / 4.0
            if law == "equally-spaced":
                x = torch.linspace(a, b, n, dtype=_F)
                sg = torch.full((n,), sd, dtype=_F)
            elif law == "uniform":
                x = torch.rand(n) * (b - a) + a
                sg = torch.zeros(n, dtype=_F)
            elif law == "log-spaced":
                x = torch.logspace(math.log10(a), math.log10(b), n, dtype=_F)
                sg = sd * x
            elif law == "exp-spaced":
                x = torch.log(torch.linspace(base[d] ** a, base[d] ** b, n, dtype=_F)) / math.log(base[d])
                sg = sd * x
            elif law in ("chebyshev", "chebyshev1"):
                x = _cheb1(a, b, n)
                sg = torch.full((n,), sd, dtype=_F)
            elif law == "chebyshev2":
                x = _cheb2(a, b, n)
                sg = torch.full((n,), sd, dtype=_F)
            else:
                raise ValueError(f"Unknown method: {law}")
            nodes.append(x[cut[d][0]:cut[d][1]])
            sigmas.append(sg[cut[d][0]:cut[d][1]])
        self._mesh = [m.flatten() for m in torch.meshgrid(*nodes, indexing="ij")]
        self._std = [m.flatten() for m in torch.meshgrid(*sigmas, indexing="ij")]
        self._fold = abs_value

    def get_examples(self):
        if not self.noisy:
            return tuple(self._mesh)
        pts = tuple(torch.normal(m, s) for m, s in zip(self._mesh, self._std))
        return tuple(p.abs() for p in pts) if self._fold else pts


class SamplerGenerator(BaseGenerator):
    """Normalises any generator's output to a list of (N, 1) columns (generators.py:1046-1057)."""

    def __init__(self, generator):
        super().__init__()
        self.generator = generator
        self.size = generator.size

    def get_examples(self):
        return [u.reshape(-1, 1) for u in _as_tuple(self.generator.get_examples())]
