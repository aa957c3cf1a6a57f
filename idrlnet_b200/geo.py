"""Geometry for the configs' shapes: `Line1D` and `Rectangle` with closed-form
signed distance, NumPy host sampling with the reference's semantics
(idrlnet/geo_utils/geo.py:80-108,191-244,316-343; geo_obj.py:30-41,109-167) and a
device sampler (`sample_interior_device`) on top of `pinn_sample_box`.

The reference's SymPy CSG algebra and the other 13 shapes are out of scope
(SURVEY.md section 2, rows 14-15)."""
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import sympy

__all__ = ["Line1D", "Rectangle", "Geometry"]


def _name(key) -> str:
    return key if isinstance(key, str) else key.name


def _ranged_sample(n: int, ranges: Dict, low_discrepancy: bool = False) -> Dict[str, np.ndarray]:
    """float -> constant column, (lo, hi) -> uniform, callable -> callable(n)."""
    out = {}
    ld_keys = []
    for key, value in ranges.items():
        k = _name(key)
        if isinstance(value, (float, int)):
            out[k] = np.ones((n, 1)) * value
        elif isinstance(value, tuple):
            assert len(value) == 2, "Tuple: length of range should be 2!"
            if low_discrepancy:
                ld_keys.append((k, value))
                continue
            out[k] = np.random.uniform(value[0], value[1], size=(n, 1))
        elif callable(value):
            out[k] = value(n)
        else:
            raise TypeError(f"range type {type(value)} not supported!")
    if ld_keys:
        pts = _halton(n, len(ld_keys))
        for j, (k, (lo, hi)) in enumerate(ld_keys):
            out[k] = lo + (hi - lo) * pts[:, j: j + 1]
    return {k: v.astype(np.float64) for k, v in out.items()}


def _halton(n: int, dims: int) -> np.ndarray:
    """Low-discrepancy points in [0,1)^dims (the reference uses its own recursive
    bisection, geo.py:370-403; any low-discrepancy set serves the same purpose)."""
    primes = [2, 3, 5, 7, 11, 13][:dims]
    out = np.zeros((n, dims))
    for j, b in enumerate(primes):
        i = np.arange(1, n + 1)
        f, r = 1.0, np.zeros(n)
        while np.any(i > 0):
            f /= b
            r += f * (i % b)
            i //= b
        out[:, j] = r
    return out


def _criteria(sieve, points: Dict[str, np.ndarray]) -> np.ndarray:
    n = len(next(iter(points.values())))
    if sieve is None or sieve is True:
        return np.ones((n, 1), dtype=bool)
    e = sympy.sympify(sieve)
    used = sorted(s.name for s in e.free_symbols)
    fn = sympy.lambdify([sympy.Symbol(k) for k in used], e, ["numpy", {"equal": np.isclose}])
    if isinstance(e, sympy.Equality):  # exact float equality is how the reference's Eq sieves behave
        lhs = sympy.lambdify([sympy.Symbol(k) for k in used], e.lhs - e.rhs, "numpy")
        res = np.isclose(lhs(*[points[k] for k in used]) + np.zeros((n, 1)), 0.0)
    else:
        res = fn(*[points[k] for k in used])
    return np.broadcast_to(np.asarray(res, dtype=bool), (n, 1))


class Geometry:
    """Axis-aligned box in `dims` (1 or 2) dimensions."""
    axes: Tuple[str, ...] = ()

    def __init__(self, lo, hi):
        self.lo = tuple(float(v) for v in lo)
        self.hi = tuple(float(v) for v in hi)
        self.bounds = {a: (l, h) for a, l, h in zip(self.axes, self.lo, self.hi)}

    # -- signed distance, positive inside ------------------------------------
    def sdf_np(self, points: Dict[str, np.ndarray]) -> np.ndarray:
        diffs = []
        for a, l, h in zip(self.axes, self.lo, self.hi):
            c = l + (h - l) / 2
            diffs.append(np.abs(points[a] - c) - (h - c))
        if len(diffs) == 1:
            return -diffs[0]
        outside = np.sqrt(sum(np.maximum(d, 0) ** 2 for d in diffs))
        inside = np.minimum(np.maximum.reduce(diffs), 0)
        return -(outside + inside)

    def _edges(self) -> List[Tuple[Callable, float]]:
        raise NotImplementedError

    def _sieve_points(self, points, sieve, tol=1e-4, sign=1.0):
        points["sdf"] = self.sdf_np(points)
        keep = np.logical_and(points["sdf"] > -tol, _criteria(sieve, points))
        if sign == -1:
            keep = np.logical_and(points["sdf"] < tol, keep)
        return {k: v[keep[:, 0], :] for k, v in points.items()}

    def sample_interior(self, density: int, bounds: Dict = None, sieve=None, param_ranges: Dict = None,
                        low_discrepancy=False) -> Dict[str, np.ndarray]:
        bounds = self.bounds if bounds is None else bounds
        bounds = {_name(k): v for k, v in bounds.items()}
        param_ranges = {} if param_ranges is None else param_ranges
        measure = np.prod([v[1] - v[0] for v in bounds.values()])
        n = int(measure * density)
        points = _ranged_sample(n, {**bounds, **param_ranges}, low_discrepancy)
        points = self._sieve_points(points, sieve, tol=0.0)
        points["area"] = np.zeros_like(points["sdf"]) + (1.0 / density)
        return points

    def sample_boundary(self, density: int, sieve=None, param_ranges: Dict = None, low_discrepancy=False):
        param_ranges = {} if param_ranges is None else param_ranges
        parts = []
        for fn, measure in self._edges():
            n = int(density * measure)
            pr = _ranged_sample(n, {"__l": (0.0, 1.0), **param_ranges}, low_discrepancy)
            pts = fn(pr.pop("__l"))
            pts["area"] = np.zeros((n, 1)) + (measure / n if n else 0.0)
            for k in param_ranges:
                pts[_name(k)] = pr[_name(k)]
            parts.append(pts)
        points = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
        return self._sieve_points(points, sieve, sign=-1, tol=1e-4)


class Line1D(Geometry):
    axes = ("x",)

    def __init__(self, point_1, point_2):
        super().__init__((point_1,), (point_2,))

    def _edges(self):
        def end(x, nx):
            return lambda l: {"x": np.zeros_like(l) + x, "normal_x": np.zeros_like(l) + nx}
        return [(end(self.lo[0], -1.0), 1.0), (end(self.hi[0], 1.0), 1.0)]


class Rectangle(Geometry):
    axes = ("x", "y")

    def __init__(self, point_1, point_2):
        super().__init__(point_1, point_2)

    def _edges(self):
        (x0, y0), (x1, y1) = self.lo, self.hi
        dx, dy = x1 - x0, y1 - y0
        z = np.zeros_like

        def mk(fx, fy, nx, ny):
            return lambda l: {"x": fx(l), "y": fy(l), "normal_x": z(l) + nx, "normal_y": z(l) + ny}
        return [
            (mk(lambda l: l * dx + x0, lambda l: z(l) + y0, 0.0, -1.0), dx),
            (mk(lambda l: z(l) + x1, lambda l: l * dy + y0, 1.0, 0.0), dy),
            (mk(lambda l: l * dx + x0, lambda l: z(l) + y1, 0.0, 1.0), dx),
            (mk(lambda l: z(l) + x0, lambda l: -l * dy + y1, -1.0, 0.0), dy),
        ]
