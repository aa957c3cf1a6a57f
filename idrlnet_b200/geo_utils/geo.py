"""Geometry kernel: parametric boundary pieces (`Edge`), signed-distance based solids with CSG
(`Geometry`), and the samplers `sample_interior` / `sample_boundary`.

Semantics follow idrlnet/geo_utils/geo.py:38-343 -- including the ORDER in which NumPy's global
RNG is consumed (a 100-point area probe per edge, then the points, ranges in insertion order), so
a seeded run draws the same point stream as the reference (tests/golden/geometry.npz).  Design
differences: every SymPy expression is compiled once and cached (the reference re-lambdifies on
every call); `Edge.rotation` is a proper rotation (the reference rotates the second coordinate with
the already rotated first one, idrlnet/geo_utils/geo.py:52-61, which moves rotated edges off the
solid); solids additionally sample on the GPU (`sample_interior_device`)."""
import copy
import itertools
import math
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import sympy
from sympy import Max, Min, Symbol, cos, sin

from .sympy_np import lambdify_np

__all__ = ["Edge", "Geometry", "Geometry1D", "Geometry2D", "Geometry3D", "device_sampling", "set_device_sampling"]


class _DeviceSampling:
    """Product-path switch for SURVEY.md 8 a7 ("geometry sampling moved on-device"): while enabled and a CUDA device is
    active, `Geometry.sample_interior` / `sample_boundary` -- the calls an unchanged example makes from its
    `SampleDomain.sampling` -- draw their points on the GPU (Philox4x32-10 keyed by (seed, stream), counter = running
    offset, so every call sees fresh points and every rank of a torch.distributed job its own stream) and return
    device tensors; nothing is sampled on the host and nothing is copied.  The point STREAM differs from NumPy's
    (statistically equivalent, tests/test_gpu_sampling.py); keep it off to reproduce a seeded reference run."""
    enabled = False
    seed = 0
    offset = 0

    @classmethod
    def next_offset(cls, n: int) -> int:
        off = cls.offset
        cls.offset += max(int(n), 1)
        return off

    @staticmethod
    def rank_world():
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        return 0, 1


def set_device_sampling(enabled: bool = True, seed: int = 0) -> None:
    """Switch every `sample_interior` / `sample_boundary` call to the GPU samplers (see `_DeviceSampling`);
    `Solver(device_sampling=True)` and IDRLNET_B200_DEVICE_SAMPLING=1 call this."""
    _DeviceSampling.enabled = bool(enabled)
    _DeviceSampling.seed = int(seed)
    _DeviceSampling.offset = 0


def device_sampling() -> bool:
    if not _DeviceSampling.enabled:
        return False
    from .. import DEVICE
    return DEVICE.type == "cuda"


def _key(k) -> str:
    return k if isinstance(k, str) else k.name


def _low_discrepancy(n: int, stack: List[Tuple[str, Tuple[float, float]]]) -> Dict[str, np.ndarray]:
    """Stratified points by recursive 2^dim-section (idrlnet/geo_utils/geo.py:370-403): the index range
    is cut into 2^dim nearly equal blocks (remainders placed by a random shuffle), block i is squeezed
    into the i-th orthant cell, recursively with halved cells."""
    dim = len(stack)
    cells = 2 ** dim
    pts = np.random.rand(n, dim)

    def split(lo: int, hi: int, corner: List[float], width: float) -> None:
        if hi - lo <= 1:
            return
        base, rem = divmod(hi - lo, cells)
        extra = (np.arange(cells - 1, 0, -1) + rem) // cells
        np.random.shuffle(extra)
        cuts = np.concatenate([[lo], (base + extra).cumsum() + lo, [hi]])
        for i in range(cells):
            a, b = cuts[i], cuts[i + 1]
            bits = [(i >> j) & 1 for j in range(dim)]
            for j in range(dim):
                pts[a:b, j] = (pts[a:b, j] - corner[j]) / 2 + corner[j] + bits[j] * width
            split(a, b, [c + width * bit for c, bit in zip(corner, bits)], width / 2)

    split(0, n, [0.0] * dim, 0.5)
    return {name: pts[:, j: j + 1] * (hi - lo) + lo for j, (name, (lo, hi)) in enumerate(stack)}


def _ranged_sample(n: int, ranges: Dict, low_discrepancy: bool = False) -> Dict[str, np.ndarray]:
    """number -> constant column, (lo, hi) -> uniform, callable -> callable(n)
    (idrlnet/geo_utils/geo.py:316-343)."""
    out: Dict[str, np.ndarray] = {}
    stack = []
    for key, value in ranges.items():
        name = _key(key)
        if isinstance(value, (float, int)):
            out[name] = np.ones((n, 1)) * value
        elif isinstance(value, tuple):
            assert len(value) == 2, "Tuple: length of range should be 2!"
            if low_discrepancy:
                stack.append((name, value))
            else:
                out[name] = np.random.uniform(value[0], value[1], size=(n, 1))
        elif callable(value):
            out[name] = value(n)
        else:
            raise TypeError(f"range type {type(value)} not supported!")
    if low_discrepancy:
        out.update(_low_discrepancy(n, stack))
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def _ranged_sample_quiet(n: int, ranges: Dict) -> Dict[str, np.ndarray]:
    """`_ranged_sample` on a private generator: does not disturb the global NumPy stream."""
    rng = np.random.default_rng(0)
    out = {}
    for key, value in ranges.items():
        if isinstance(value, (float, int)):
            out[_key(key)] = np.ones((n, 1)) * value
        elif isinstance(value, tuple):
            out[_key(key)] = rng.uniform(value[0], value[1], size=(n, 1))
        else:
            out[_key(key)] = np.asarray(value(n), dtype=np.float64)
    return out


class Edge:
    """A boundary piece: coordinates and normals as functions of its parameters, with the
    parameter ranges and its measure (`area`, possibly symbolic in shape parameters)."""

    def __init__(self, functions: Dict, ranges: Dict, area):
        self.functions = dict(functions)
        self.ranges = dict(ranges)
        self.area = area

    @property
    def axes(self) -> List[str]:
        return [k for k in self.functions if not k.startswith("normal")]

    def rotation(self, angle: float, axis: str = "z"):
        assert len(self.axes) > 1, "Cannot rotate a object with dim<2"
        a, b = [k for k in self.axes if k != axis][:2]
        c, s = cos(angle), sin(angle)
        f = self.functions
        for u, v in ((a, b), ("normal_" + a, "normal_" + b)):
            f[u], f[v] = c * f[u] - s * f[v], s * f[u] + c * f[v]
        return self

    def scaling(self, scale: float):
        for k in self.axes:
            self.functions[k] = self.functions[k] * scale
        self.area = scale ** (len(self.axes) - 1) * self.area
        return self

    def translation(self, direction):
        assert len(direction) == len(self.axes), "Moving direction must have the save dimension with the object"
        for k, d in zip(self.axes, direction):
            self.functions[k] = self.functions[k] + d
        return self

    def inverted(self) -> "Edge":
        return Edge({k: (-v if k.startswith("normal_") else v) for k, v in self.functions.items()}, self.ranges, self.area)

    def sample(self, density: int, param_ranges=None, low_discrepancy=False) -> Dict[str, np.ndarray]:
        param_ranges = {} if param_ranges is None else param_ranges
        ranges = {**self.ranges, **param_ranges}
        names = [_key(k) for k in ranges]
        area_fn = lambdify_np(self.area, names)
        probe = _ranged_sample(100, ranges)                      # expected measure over the parameter ranges
        n = int(density * np.mean(area_fn(**probe)))
        params = _ranged_sample(n, ranges, low_discrepancy)
        out = {"area": area_fn(**params) / n} if n else {"area": np.zeros((0, 1))}
        for k, fn in self.functions.items():
            out[k] = lambdify_np(fn, names)(**params)
        for k in param_ranges:
            out[_key(k)] = params[_key(k)]
        return out


def _rotate_bounds(bx: Tuple, by: Tuple, angle: float):
    try:
        c, s = math.cos(angle), math.sin(angle)
    except TypeError:  # symbolic angle: the bounding square of a 45 degree turn covers every case
        c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    xs, ys = zip(*[(c * x - s * y, s * x + c * y) for x, y in itertools.product(bx, by)])
    return (min(xs), max(xs)), (min(ys), max(ys))


def _combine(kind: str, f, g):
    """max / min / negation of signed-distance fields that are SymPy expressions or callables."""
    symbolic = all(isinstance(h, sympy.Basic) or isinstance(h, (int, float)) for h in (f, g) if h is not None)
    if kind == "neg":
        return -f if symbolic else (lambda **x: -lambdify_np(f, x)(**x))
    if symbolic:
        return Max(f, g) if kind == "max" else Min(f, g)
    op = np.maximum if kind == "max" else np.minimum
    return lambda **x: op(lambdify_np(f, x)(**x), lambdify_np(g, x)(**x))


def _program(tree) -> List[Tuple[str, Tuple[float, ...]]]:
    """Post-order (op, params) list of a CSG tree: ("prim", name, params) | (op, left[, right])."""
    if tree[0] == "prim":
        return [(tree[1], tuple(float(v) for v in tree[2]))]
    out = []
    for sub in tree[1:]:
        out += _program(sub)
    return out + [(tree[0], ())]


_CONST_COLUMNS: Dict = {}


def _const_column(m: int, value: float, dev):
    """[m, 1] float32 device column filled with `value`, created once per (m, value, device): constant parameter and
    `area` columns are read-only downstream, so every sampling call can hand out the same tensor (no fill launch)."""
    import torch
    key = (int(m), float(value), str(dev))
    col = _CONST_COLUMNS.get(key)
    if col is None:
        if len(_CONST_COLUMNS) > 64:
            _CONST_COLUMNS.clear()
        col = _CONST_COLUMNS[key] = torch.full((m, 1), float(value), device=dev, dtype=torch.float32)
    return col


def _box_fills_bounds(tree, bounds: Dict) -> bool:
    """True when the primitive ("interval" | "rect" | "box") coincides with the box it is sampled over."""
    kind, p = tree[1], [float(v) for v in tree[2]]
    b = [(float(lo), float(hi)) for lo, hi in bounds.values()]
    if kind == "interval":
        return len(b) == 1 and abs(p[0] - b[0][0]) < 1e-12 and abs(p[1] - b[0][1]) < 1e-12
    d = len(b)
    if (kind == "rect" and d != 2) or (kind == "box" and d != 3):
        return False
    return all(abs(p[i] - b[i][0]) < 1e-12 and abs(p[d + i] - b[i][1]) < 1e-12 for i in range(d))


class Geometry:
    """A solid: `sdf` (positive inside), bounding box `bounds` and boundary `edges`."""
    edges: List[Edge] = None
    bounds: Dict = None
    sdf = None
    # CSG tree over numeric primitives, kept while the solid is built from untransformed, non-symbolic
    # shapes: lets `sample_interior_device` run the hand-written CSG kernel (pinn_sample_csg)
    _tree = None

    def _join(self, op: str, g: "Geometry", *others: "Geometry") -> None:
        parts = (self,) + others
        g._tree = (op,) + tuple(p._tree for p in parts) if all(p._tree is not None for p in parts) else None

    @property
    def axes(self) -> List[str]:
        return self.edges[0].axes

    def _blank(self, other: Optional["Geometry"] = None) -> "Geometry":
        for base in (Geometry1D, Geometry2D, Geometry3D):
            if isinstance(self, base):
                assert other is None or isinstance(other, base), "CSG needs solids of the same dimension"
                return base()
        raise TypeError("CSG needs a Geometry1D/2D/3D")

    # -- rigid motions and scaling ------------------------------------------------------------
    def _need_symbolic_sdf(self):
        if not isinstance(self.sdf, sympy.Basic):
            raise NotImplementedError("this solid's sdf is a Python callable and cannot be transformed")

    def translation(self, direction: Union[List, Tuple]) -> "Geometry":
        assert len(direction) == len(self.axes)
        self._need_symbolic_sdf()
        self._tree = None
        for e in self.edges:
            e.translation(direction)
        self.sdf = self.sdf.subs([(Symbol(a), Symbol(a) - d) for a, d in zip(self.axes, direction)])
        self.bounds = {**self.bounds, **{a: (self.bounds[a][0] + d, self.bounds[a][1] + d) for a, d in zip(self.axes, direction)}}
        return self

    def rotation(self, angle: float, axis: str = "z", center=None) -> "Geometry":
        self._need_symbolic_sdf()
        self._tree = None
        if center is not None:
            self.translation([-c for c in center])
        for e in self.edges:
            e.rotation(angle, axis)
        a, b = [k for k in self.axes if k != axis][:2]
        sa, sb, ta, tb = Symbol(a), Symbol(b), Symbol("_rot_a"), Symbol("_rot_b")
        self.sdf = self.sdf.subs({sa: cos(angle) * ta + sin(angle) * tb, sb: -sin(angle) * ta + cos(angle) * tb},
                                 simultaneous=True).subs({ta: sa, tb: sb}, simultaneous=True)
        self.bounds = dict(self.bounds)
        self.bounds[a], self.bounds[b] = _rotate_bounds(self.bounds[a], self.bounds[b], angle)
        if center is not None:
            self.translation(center)
        return self

    def scaling(self, scale: float, center: Tuple = None) -> "Geometry":
        assert scale > 0, "scaling must be positive"
        self._need_symbolic_sdf()
        self._tree = None
        if center is not None:
            self.translation(tuple(-c for c in center))
        for e in self.edges:
            e.scaling(scale)
        self.sdf = scale * self.sdf.subs({Symbol(a): Symbol(a) / scale for a in self.axes}, simultaneous=True)
        self.bounds = {**self.bounds, **{a: (self.bounds[a][0] * scale, self.bounds[a][1] * scale) for a in self.axes}}
        if center is not None:
            self.translation(center)
        return self

    def duplicate(self) -> "Geometry":
        return copy.deepcopy(self)

    # -- CSG ----------------------------------------------------------------------------------------
    def __add__(self, other: "Geometry") -> "Geometry":
        g = self._blank(other)
        g.edges = self.edges + other.edges
        g.sdf = _combine("max", self.sdf, other.sdf)
        g.bounds = {k: (min(v[0], other.bounds[k][0]), max(v[1], other.bounds[k][1])) for k, v in self.bounds.items()}
        self._join("union", g, other)
        return g

    def __sub__(self, other: "Geometry") -> "Geometry":
        g = self._blank(other)
        g.edges = self.edges + [e.inverted() for e in other.edges]
        g.sdf = _combine("min", self.sdf, _combine("neg", other.sdf, None))
        g.bounds = dict(self.bounds)
        self._join("difference", g, other)
        return g

    def __and__(self, other: "Geometry") -> "Geometry":
        g = self._blank(other)
        g.edges = self.edges + other.edges
        g.sdf = _combine("min", self.sdf, other.sdf)
        g.bounds = {k: (max(v[0], other.bounds[k][0]), min(v[1], other.bounds[k][1])) for k, v in self.bounds.items()}
        self._join("intersection", g, other)
        return g

    def __invert__(self) -> "Geometry":
        g = self._blank()
        g.edges = [e.inverted() for e in self.edges]
        g.sdf = _combine("neg", self.sdf, None)
        g.bounds = {k: (-float("inf"), float("inf")) for k in self.bounds}
        self._join("complement", g)
        return g

    # -- sampling ---------------------------------------------------------------------------------
    def _sieve_points(self, points: Dict[str, np.ndarray], sieve, tol=1e-4, sign=1.0) -> Dict[str, np.ndarray]:
        names = list(points.keys())
        points["sdf"] = lambdify_np(self.sdf, names)(**{k: points[k] for k in names})
        keep = np.logical_and(points["sdf"] > -tol, lambdify_np(True if sieve is None else sieve, list(points))(**points))
        if sign == -1:
            keep = np.logical_and(points["sdf"] < tol, keep)
        return {k: v[keep[:, 0], :] for k, v in points.items()}

    def sample_boundary(self, density: int, sieve=None, param_ranges: Dict = None, low_discrepancy=False):
        if device_sampling() and not low_discrepancy:
            return self.sample_boundary_device(density, seed=_DeviceSampling.seed, offset=None, sieve=sieve, param_ranges=param_ranges)
        param_ranges = {} if param_ranges is None else param_ranges
        parts = [e.sample(density, param_ranges, low_discrepancy) for e in self.edges]
        points = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
        return self._sieve_points(points, sieve, sign=-1, tol=1e-4)

    def sample_interior(self, density: int, bounds: Dict = None, sieve=None, param_ranges: Dict = None,
                        low_discrepancy=False) -> Dict[str, np.ndarray]:
        if device_sampling() and not low_discrepancy:
            return self.sample_interior_device(density, seed=_DeviceSampling.seed, offset=None, bounds=bounds, sieve=sieve,
                                               param_ranges=param_ranges)
        bounds = self.bounds if bounds is None else bounds
        bounds = {_key(k): v for k, v in bounds.items()}
        param_ranges = {} if param_ranges is None else param_ranges
        measure = np.prod([hi - lo for lo, hi in bounds.values()])
        n = int(measure * density)
        points = _ranged_sample(n, {**bounds, **param_ranges}, low_discrepancy)
        points = self._sieve_points(points, sieve, tol=0.0)
        points["area"] = np.zeros_like(points["sdf"]) + (1.0 / density)
        return points

    def sample_interior_device(self, density: int, seed: int = 0, offset: Optional[int] = 0, bounds: Dict = None, sieve=None,
                               device=None, param_ranges: Dict = None):
        """`sample_interior` on the GPU.  Candidates are the Philox points of pinn_sample_box in the bounding
        box; a solid built from numeric primitives and CSG operators is evaluated and compacted by the
        hand-written kernel (pinn_sample_csg: sdf program, ordered compaction, deterministic); a box-shaped
        primitive sampled over its own bounding box needs no rejection at all (one launch, no host sync); anything
        else (symbolic parameters, transformed solids, a `sieve`) falls back to the same SymPy expressions lowered
        to torch.  `param_ranges` columns (time, design parameters: idrlnet/geo_utils/geo.py:233-240) are drawn for the
        surviving points on their own Philox stream.  `offset=None` takes the next offset of the process-wide counter
        (fresh points on every call); under torch.distributed the stream id carries the rank.  Returns a dict of
        [n, 1] float32 device tensors (coordinates, parameters, `sdf`, `area`)."""
        import ctypes as C

        import torch

        from .. import native as nv
        from ..torch_util import torch_lambdify
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        bounds = {_key(k): v for k, v in (self.bounds if bounds is None else bounds).items()}
        param_ranges = {} if param_ranges is None else {_key(k): v for k, v in param_ranges.items()}
        names = list(bounds)
        n = int(np.prod([hi - lo for lo, hi in bounds.values()]) * density)
        rank, world = _DeviceSampling.rank_world()
        if offset is None:
            offset = _DeviceSampling.next_offset(n)
        sid = 64 * rank  # Philox stream ids 64*rank .. 64*rank + 63 belong to this rank
        cols = [torch.empty(n, device=dev, dtype=torch.float32) for _ in names]
        lo = (C.c_float * len(names))(*[float(bounds[k][0]) for k in names])
        hi = (C.c_float * len(names))(*[float(bounds[k][1]) for k in names])
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib = nv.lib()

        def with_params(out, m):
            drawn = [(k, v) for k, v in param_ranges.items() if isinstance(v, tuple)]
            for g0 in range(0, len(drawn), 4):
                grp = drawn[g0:g0 + 4]
                pc = [torch.empty(m, device=dev, dtype=torch.float32) for _ in grp]
                plo = (C.c_float * len(grp))(*[float(v[0]) for _, v in grp])
                phi = (C.c_float * len(grp))(*[float(v[1]) for _, v in grp])
                if m:
                    nv.check(lib.pinn_sample_box(nv.ptr_array([c.data_ptr() for c in pc]), len(grp), plo, phi, m, seed,
                                                 sid + 32 + g0 // 4, offset, None, 0, stream))
                for (k, _), c in zip(grp, pc):
                    out[k] = c.reshape(-1, 1)
            for k, v in param_ranges.items():
                if isinstance(v, (int, float)):
                    out[k] = _const_column(m, float(v), dev)
                elif not isinstance(v, tuple):
                    out[k] = torch.as_tensor(np.asarray(v(m)), dtype=torch.float32, device=dev).reshape(-1, 1)
            out["area"] = _const_column(m, 1.0 / density, dev)
            return out

        tree_ok = self._tree is not None and sieve is None and names == list(self.axes)
        if tree_ok and self._tree[0] == "prim" and self._tree[1] in ("interval", "rect", "box") and _box_fills_bounds(self._tree, bounds):
            # the solid IS the sampling box: every candidate is interior (a coordinate exactly on a face, probability 2^-24
            # per draw, is kept: a null set) -> one launch, the count is known on the host, no synchronisation
            sdf = torch.empty(n, device=dev, dtype=torch.float32)
            if n:
                nv.check(lib.pinn_sample_box(nv.ptr_array([c.data_ptr() for c in cols]), len(names), lo, hi, n, seed, sid,
                                             offset, sdf.data_ptr(), len(names), stream))
            out = {k: c.reshape(-1, 1) for k, c in zip(names, cols)}
            out["sdf"] = sdf.reshape(-1, 1)
            return with_params(out, n)
        if tree_ok:
            prog = _program(self._tree)
            ops = (nv.PinnShapeOp * len(prog))()
            for o, (name, params) in zip(ops, prog):
                o.op = nv.SHAPE_OPS[name]
                for i, v in enumerate(params):
                    o.p[i] = v
            sdf = torch.empty(n, device=dev, dtype=torch.float32)
            count = torch.zeros(1, device=dev, dtype=torch.int64)
            ws_bytes = lib.pinn_sample_csg_workspace_bytes(n)
            ws = torch.empty(max(ws_bytes, 1), device=dev, dtype=torch.uint8)
            nv.check(lib.pinn_sample_csg(nv.ptr_array([c.data_ptr() for c in cols]), len(names), lo, hi, n, n, seed, sid, offset,
                                         ops, len(prog), sdf.data_ptr(), count.data_ptr(), ws.data_ptr(), ws_bytes, stream))
            m = int(count.item())
            out = {k: c[:m].reshape(-1, 1) for k, c in zip(names, cols)}
            out["sdf"] = sdf[:m].reshape(-1, 1)
            return with_params(out, m)
        if n:
            nv.check(lib.pinn_sample_box(nv.ptr_array([c.data_ptr() for c in cols]), len(names), lo, hi, n, seed, sid,
                                         offset, None, 0, stream))
        pts = with_params({k: c.reshape(-1, 1) for k, c in zip(names, cols)}, n)
        area = pts.pop("area")
        if not isinstance(self.sdf, sympy.Basic):
            raise NotImplementedError("device sampling needs a symbolic sdf")
        keys = list(pts)
        sdf = torch_lambdify(keys, self.sdf)(**pts)
        sdf = sdf if isinstance(sdf, torch.Tensor) and sdf.shape == pts[names[0]].shape else torch.zeros_like(pts[names[0]]) + sdf
        keep = sdf > 0
        if sieve is not None:
            keep = keep & torch_lambdify(keys + ["sdf"], sieve)(**pts, sdf=sdf)
        mask = keep[:, 0]
        out = {k: v[mask] for k, v in pts.items()}
        out["sdf"] = sdf[mask]
        out["area"] = area[mask]
        return out

    def sample_boundary_device(self, density: int, seed: int = 0, offset: Optional[int] = 0, sieve=None, param_ranges: Dict = None,
                               device=None):
        """`sample_boundary` on the GPU: per edge, `int(density * measure)` Philox parameter draws
        (pinn_sample_box), the edge's coordinate / normal expressions and the solid's signed distance lowered
        to torch, then the reference's sieve (|sdf| < 1e-4 and the user's criterion).  Returns [n, 1] float32
        device tensors with the same keys as `sample_boundary`."""
        import ctypes as C

        import torch

        from .. import native as nv
        from ..torch_util import torch_lambdify
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        param_ranges = {} if param_ranges is None else param_ranges
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
        lib = nv.lib()
        rank, _ = _DeviceSampling.rank_world()
        auto = offset is None
        if dev.type == "cuda":
            fast = self._sample_boundary_programs(density, seed, offset, sieve, param_ranges, dev, stream, rank)
            if fast is not None:
                return fast
        parts = []
        for e_idx, edge in enumerate(self.edges):
            ranges = {**edge.ranges, **param_ranges}
            names = [_key(k) for k in ranges]
            area_fn = lambdify_np(edge.area, names)
            n = int(density * np.mean(area_fn(**_ranged_sample_quiet(100, ranges))))   # host probe of the measure only
            if auto:
                offset = _DeviceSampling.next_offset(n)
            params = {}
            drawn = [(k, v) for k, v in ranges.items() if isinstance(v, tuple)]
            for g0 in range(0, len(drawn), 4):  # the box sampler draws up to 4 columns per call
                grp = drawn[g0:g0 + 4]
                cols = [torch.empty(n, device=dev, dtype=torch.float32) for _ in grp]
                lo = (C.c_float * len(grp))(*[float(v[0]) for _, v in grp])
                hi = (C.c_float * len(grp))(*[float(v[1]) for _, v in grp])
                if n:
                    nv.check(lib.pinn_sample_box(nv.ptr_array([c.data_ptr() for c in cols]), len(grp), lo, hi, n, seed,
                                                 (rank << 20) + 1024 + e_idx * 16 + g0 // 4, offset, None, 0, stream))
                for (k, _), c in zip(grp, cols):
                    params[_key(k)] = c.reshape(-1, 1)
            for k, v in ranges.items():
                if isinstance(v, (int, float)):
                    params[_key(k)] = torch.full((n, 1), float(v), device=dev)
                elif not isinstance(v, tuple):
                    params[_key(k)] = torch.as_tensor(np.asarray(v(n)), dtype=torch.float32, device=dev).reshape(-1, 1)
            like = torch.zeros((n, 1), device=dev)

            def column(expr):
                out = torch_lambdify(names, expr)(**{k: params[k] for k in names})
                return out + like if isinstance(out, torch.Tensor) else like + float(out)
            pts = {"area": column(edge.area) / max(n, 1)}
            for k, fn in edge.functions.items():
                pts[k] = column(fn)
            for k in param_ranges:
                pts[_key(k)] = params[_key(k)]
            parts.append(pts)
        points = {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}
        if not isinstance(self.sdf, sympy.Basic):
            raise NotImplementedError("device sampling needs a symbolic sdf")
        keys = list(points)
        sdf = torch_lambdify(keys, self.sdf)(**points)
        points["sdf"] = sdf + torch.zeros_like(points[keys[0]]) if isinstance(sdf, torch.Tensor) else torch.zeros_like(points[keys[0]]) + float(sdf)
        if sieve is None and self._tree is not None and self._tree[0] == "prim":
            # an untransformed primitive without a sieve: every edge sample lies on the solid's own boundary (|sdf| <= 1e-4
            # by construction), nothing is dropped -> no mask, no device-to-host synchronisation
            return points
        keep = (points["sdf"] > -1e-4) & (points["sdf"] < 1e-4)
        if sieve is not None:
            keep = keep & torch_lambdify(list(points), sieve)(**points)
        mask = keep[:, 0]
        return {k: v[mask] for k, v in points.items()}


    # -- boundary sampling by compiled edge programs (csrc/geometry.cu: pinn_sample_program) ----------------
    def _edge_programs(self, density: int, param_ranges: Dict):
        """Per edge: (n, packed programs, parameter bounds) -- compiled once per (solid state, density, ranges) and cached.
        Output order = the keys `sample_boundary` returns: area, edge functions, parameter columns, sdf."""
        from . import expr_prog as xp
        # identity of the (immutable) expression objects: a moved / rotated / scaled solid holds new ones.  The entry pins
        # the objects so that an id cannot be recycled while it is a key.
        key = (density, tuple((_key(k), v if not callable(v) else id(v)) for k, v in param_ranges.items()), id(self.sdf),
               tuple(id(f) for e in self.edges for f in e.functions.values()))
        cache = self.__dict__.setdefault("_prog_cache", {})
        if key in cache:
            return cache[key][0]
        pins = (self.sdf, [f for e in self.edges for f in e.functions.values()])
        entry = None
        try:
            if not isinstance(self.sdf, sympy.Basic):
                raise xp.Unsupported("callable sdf")
            plan, keys = [], None
            for edge in self.edges:
                ranges = {**edge.ranges, **param_ranges}
                names = [_key(k) for k in ranges]
                drawn = [(_key(k), v) for k, v in ranges.items() if isinstance(v, tuple)]
                consts = {_key(k): float(v) for k, v in ranges.items() if isinstance(v, (int, float))}
                if any(callable(v) for v in ranges.values()) or len(drawn) > nv_expr_max_vars():
                    raise xp.Unsupported("callable range / too many parameters")
                var_index = {k: i for i, (k, _) in enumerate(drawn)}
                area_fn = lambdify_np(edge.area, names)
                n = int(density * np.mean(area_fn(**_ranged_sample_quiet(100, ranges))))   # host probe of the measure only
                out_keys = ["area"] + list(edge.functions) + [_key(k) for k in param_ranges] + ["sdf"]
                if keys is None:
                    keys = out_keys
                elif keys != out_keys:
                    raise xp.Unsupported("edges with different columns")
                progs = [xp.compile_expr(sympy.sympify(edge.area) / max(n, 1), var_index, consts)]
                progs += [xp.compile_expr(fn, var_index, consts) for fn in edge.functions.values()]
                progs += [xp.compile_expr(Symbol(_key(k)), var_index, consts) for k in param_ranges]
                on_edge = {Symbol(a): sympy.sympify(edge.functions[a]) for a in edge.axes}
                progs.append(xp.compile_expr(self.sdf.subs(on_edge, simultaneous=True), var_index, consts))
                ops, begin = xp.pack_programs(progs)
                import ctypes as C
                lo = (C.c_float * max(len(drawn), 1))(*[float(v[0]) for _, v in drawn])
                hi = (C.c_float * max(len(drawn), 1))(*[float(v[1]) for _, v in drawn])
                plan.append((n, ops, begin, lo, hi, len(drawn)))
            entry = (keys, plan)
        except xp.Unsupported:
            entry = None
        cache[key] = (entry, pins)
        return entry

    def _sample_boundary_programs(self, density, seed, offset, sieve, param_ranges, dev, stream, rank):
        """One launch per edge writes every column of its samples straight into the concatenated output; None when an
        expression has no device op (the caller then lowers the same SymPy expressions to torch)."""
        import torch

        from .. import native as nv
        from ..torch_util import torch_lambdify
        entry = self._edge_programs(density, param_ranges)
        if entry is None:
            return None
        keys, plan = entry
        total = sum(p[0] for p in plan)
        buf = torch.empty((len(keys), max(total, 1)), device=dev, dtype=torch.float32)
        base, row, start = buf.data_ptr(), buf.stride(0) * 4, 0
        lib = nv.lib()
        for e_idx, (n, ops, begin, lo, hi, n_vars) in enumerate(plan):
            off = _DeviceSampling.next_offset(n) if offset is None else offset
            if n:
                ptrs = nv.ptr_array([base + r * row + start * 4 for r in range(len(keys))])
                nv.check(lib.pinn_sample_program(ptrs, len(keys), ops, begin, lo, hi, n_vars, n, seed,
                                                 (rank << 20) + 1024 + e_idx * 16, off, stream))
            start += n
        points = {k: buf[i, :total].unsqueeze(1) for i, k in enumerate(keys)}
        if sieve is None and self._tree is not None and self._tree[0] == "prim":
            return points  # every edge sample of an untransformed primitive lies on its own boundary: nothing to drop, no sync
        keep = (points["sdf"] > -1e-4) & (points["sdf"] < 1e-4)
        if sieve is not None:
            keep = keep & torch_lambdify(list(points), sieve)(**points)
        mask = keep[:, 0]
        return {k: v[mask] for k, v in points.items()}


def nv_expr_max_vars() -> int:
    from .. import native as nv
    return nv.EXPR_MAX_VARS


class Geometry1D(Geometry):
    pass


class Geometry2D(Geometry):
    pass


class Geometry3D(Geometry):
    pass
