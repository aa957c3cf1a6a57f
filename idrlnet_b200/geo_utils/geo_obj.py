"""Concrete solids (the shapes of idrlnet/geo_utils/geo_obj.py:30-762).  Shape parameters may be
numbers or SymPy symbols that are later sampled through `param_ranges`.  Edge order, parameter
names and range order are part of the contract: they fix how the global NumPy RNG is consumed."""
import math
from typing import Dict, Sequence, Tuple

import numpy as np
from sympy import Abs, Heaviside, Max, Min, Symbol, asin, cos, pi, sign, sin, sqrt, symbols

from .geo import Edge, Geometry1D, Geometry2D, Geometry3D

__all__ = ["Line1D", "Line", "Tube2D", "Rectangle", "Circle", "Heart", "Triangle", "Polygon", "Plane", "Tube3D",
           "Tube", "CircularTube", "Box", "Sphere", "Cylinder"]

_XYZ = symbols("x y z")


def _box_bounds(p1, p2, axes="xyz") -> Dict[str, Tuple]:
    return {a: (min(u, v), max(u, v)) for a, u, v in zip(axes, p1, p2)}


def _slab_sdf(coords: Sequence, centers: Sequence, halves: Sequence):
    """Signed distance (positive inside) to the axis-aligned slab/box |c_i - center_i| <= half_i."""
    d = [Abs(c - m) - h for c, m, h in zip(coords, centers, halves)]
    outside = sqrt(sum(Max(v, 0) ** 2 for v in d))
    inside = Min(Max(*d), 0) if len(d) > 1 else Min(d[0], 0)
    return -(outside + inside)


def _numeric(*values) -> bool:
    return all(isinstance(v, (int, float)) for v in values)


def _normals(axes: str, **nonzero) -> Dict[str, float]:
    return {"normal_" + a: nonzero.get(a, 0) for a in axes}


class Line1D(Geometry1D):
    def __init__(self, point_1, point_2):
        x = _XYZ[0]
        ranges = {Symbol("none"): (0, 1)}
        self.edges = [Edge({"x": point_1, "normal_x": -1}, ranges, 1.0), Edge({"x": point_2, "normal_x": 1}, ranges, 1.0)]
        half = (point_2 - point_1) / 2
        self.sdf = half - Abs(x - (point_1 + half))
        self.bounds = {"x": (point_1, point_2)}
        if _numeric(point_1, point_2):
            self._tree = ("prim", "interval", (point_1, point_2, 0))


class Line(Geometry2D):
    """Segment from point_1 to point_2; `normal` = +-1 picks the side the normal points to."""

    def __init__(self, point_1, point_2, normal=1):
        x, y = _XYZ[:2]
        l = Symbol("l")
        dx, dy = point_2[0] - point_1[0], point_2[1] - point_1[1]
        length = math.sqrt(dx ** 2 + dy ** 2)
        self.edges = [Edge({"x": point_1[0] + l * dx, "y": point_1[1] + l * dy,
                            "normal_x": -dy * normal / length, "normal_y": dx * normal / length}, {l: (0, 1)}, length)]
        self.sdf = ((x - point_1[0]) * dy - (y - point_1[1]) * dx) / length
        self.bounds = _box_bounds(point_1, point_2, "xy")


class Tube2D(Geometry2D):
    """Channel between y = point_1[1] and y = point_2[1] (only those two walls are boundary)."""

    def __init__(self, point_1, point_2):
        y = _XYZ[1]
        l = Symbol("l")
        dx, dy = point_2[0] - point_1[0], point_2[1] - point_1[1]
        self.edges = [Edge({"x": l * dx + point_1[0], "y": wall, **_normals("xy", y=n)}, {l: (0, 1)}, dx)
                      for wall, n in ((point_1[1], -1), (point_2[1], 1))]
        cy = point_1[1] + dy / 2
        self.sdf = _slab_sdf([y], [cy], [point_2[1] - cy])
        self.bounds = _box_bounds(point_1, point_2, "xy")
        if _numeric(*point_1, *point_2):
            self._tree = ("prim", "channel2d", (point_1[1], point_2[1]))


class Rectangle(Geometry2D):
    def __init__(self, point_1, point_2):
        x, y = _XYZ[:2]
        l = Symbol("l")
        (x0, y0), (x1, y1) = point_1, point_2
        dx, dy = x1 - x0, y1 - y0
        sides = [  # counter-clockwise from the bottom side
            ({"x": l * dx + x0, "y": y0}, dict(y=-1), dx),
            ({"x": x1, "y": l * dy + y0}, dict(x=1), dy),
            ({"x": l * dx + x0, "y": y1}, dict(y=1), dx),
            ({"x": x0, "y": -l * dy + y1}, dict(x=-1), dy),
        ]
        self.edges = [Edge({**pos, **_normals("xy", **n)}, {l: (0, 1)}, length) for pos, n, length in sides]
        cx, cy = x0 + dx / 2, y0 + dy / 2
        self.sdf = _slab_sdf([x, y], [cx, cy], [x1 - cx, y1 - cy])
        self.bounds = _box_bounds(point_1, point_2, "xy")
        if _numeric(x0, y0, x1, y1):
            self._tree = ("prim", "rect", (x0, y0, x1, y1))


class Circle(Geometry2D):
    def __init__(self, center, radius):
        x, y = _XYZ[:2]
        theta = Symbol("theta")
        self.edges = [Edge({"x": center[0] + radius * cos(theta), "y": center[1] + radius * sin(theta),
                            "normal_x": 1 * cos(theta), "normal_y": 1 * sin(theta)},
                           {theta: (0, 2 * math.pi)}, 2 * pi * radius)]
        self.sdf = radius - sqrt((x - center[0]) ** 2 + (y - center[1]) ** 2)
        self.bounds = {"x": (center[0] - radius, center[0] + radius), "y": (center[1] - radius, center[1] + radius)}
        if _numeric(*center, radius):
            self._tree = ("prim", "circle", (center[0], center[1], radius))


class Heart(Geometry2D):
    """Two straight flanks below two half-circle lobes (idrlnet/geo_utils/geo_obj.py:193-276)."""

    def __init__(self, center=(0, 0.5), radius=0.5):
        c1, c2 = center
        arc, lin = Symbol("t"), Symbol("theta")   # range order (arc first) fixes the RNG stream
        ranges = {arc: (0, math.pi), lin: (0, 1)}
        r2 = math.sqrt(2)
        edges = []
        for side in (-1, 1):
            edges.append(Edge({"x": c1 + side * lin * radius, "y": c2 - (1 - lin) * radius,
                               "normal_x": float(side), "normal_y": -1.0}, ranges, r2 * radius))
        for side, start in ((-1, 5), (1, 3)):
            ang = math.pi / 4 * start - arc
            edges.append(Edge({"x": c1 + side * radius / 2 + radius / r2 * cos(ang), "y": c2 + radius / 2 + radius / r2 * sin(ang),
                               "normal_x": cos(ang), "normal_y": sin(ang)}, ranges, r2 * radius * math.pi))
        self.edges = edges
        x, y = _XYZ[:2]
        u, v = (x - c1) * 0.5 / radius, (y - c2) * 0.5 / radius + 0.5
        upper = Heaviside(Abs(u) + v - 1)
        lobe = sqrt((Abs(u) - 0.25) ** 2 + (v - 0.75) ** 2) - r2 / 4
        m = 0.5 * Max(Abs(u) + v, 0)
        flank = sign(Abs(u) - v) * Min(sqrt(Abs(u) ** 2 + (v - 1) ** 2), sqrt((Abs(u) - m) ** 2 + (v - m) ** 2))
        self.sdf = (-upper * lobe - (1 - upper) * flank) * radius * 2
        w = 0.5 * radius + 0.5 * r2 * radius
        self.bounds = {"x": (c1 - w, c1 + w), "y": (c2 - radius, c2 + w)}


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1]


def _cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


class Triangle(Geometry2D):
    def __init__(self, p0, p1, p2):
        x, y = _XYZ[:2]
        t = Symbol("t")
        P = [p0, p1, p2]
        e = [(P[(i + 1) % 3][0] - P[i][0], P[(i + 1) % 3][1] - P[i][1]) for i in range(3)]   # e0 = p1-p0, e1 = p2-p1, e2 = p0-p2
        v = [(x - q[0], y - q[1]) for q in P]
        s = sign(_cross(e[0], e[2]))
        d2, side = [], []
        for ei, vi in zip(e, v):
            h = Max(Min(_dot(vi, ei) / _dot(ei, ei), 1), 0)
            pq = (vi[0] - ei[0] * h, vi[1] - ei[1] * h)
            d2.append(_dot(pq, pq))
            side.append(s * _cross(vi, ei))
        self.sdf = sqrt(Min(*d2)) * sign(Min(*side))
        outward = -sign(_cross(e[0], e[1]))
        self.edges = []
        for a, b in ((1, 0), (2, 1), (0, 2)):           # each side walked from P[a] to P[b]
            length = sqrt((P[b][0] - P[a][0]) ** 2 + (P[b][1] - P[a][1]) ** 2)
            self.edges.append(Edge({"x": P[a][0] + t * (P[b][0] - P[a][0]), "y": P[a][1] + t * (P[b][1] - P[a][1]),
                                    "normal_x": (P[b][1] - P[a][1]) / length * outward,
                                    "normal_y": (P[a][0] - P[b][0]) / length * outward}, {t: (0, 1)}, length))
        xs, ys = zip(p0, p1, p2)
        self.bounds = {"x": (min(xs), max(xs)), "y": (min(ys), max(ys))}


class Polygon(Geometry2D):
    """Simple polygon (counter-clockwise vertices); numeric signed distance by the winding rule."""

    def __init__(self, points):
        verts = [np.asarray(p, dtype=np.float64) for p in points]
        t = Symbol("t")

        def sdf(x: np.ndarray, y: np.ndarray, **kwargs):
            p = np.concatenate([x, y], axis=1)
            d2 = ((verts[0] - p) ** 2).sum(axis=1, keepdims=True)
            flip = np.ones_like(x)
            for i in range(len(verts)):
                a, b = verts[i], verts[i - 1]
                e, w = b - a, p - a
                nearest = w - e * np.clip((w * e).sum(axis=1, keepdims=True) / (e * e).sum(), 0, 1)
                d2 = np.minimum(d2, (nearest ** 2).sum(axis=1, keepdims=True))
                c = np.stack([p[:, 1:] >= a[1], p[:, 1:] < b[1], e[0] * w[:, 1:] > e[1] * w[:, :1]])
                flip[np.logical_or(c.all(axis=0), np.logical_not(c).all(axis=0))] *= -1
            return -np.sqrt(d2) * flip

        self.sdf = sdf
        self.edges = []
        for i in range(len(points)):
            (ax, ay), (bx, by) = points[i - 1], points[i]
            length = math.sqrt((ax - bx) ** 2 + (ay - by) ** 2)
            self.edges.append(Edge({"x": ax - t * (ax - bx), "y": ay - t * (ay - by),
                                    "normal_x": (by - ay) / length, "normal_y": (ax - bx) / length}, {t: (0, 1)}, length))
        xs, ys = zip(*points)
        self.bounds = {"x": (min(xs), max(xs)), "y": (min(ys), max(ys))}


def _centre_sides(p1, p2):
    return [p1[i] + (p2[i] - p1[i]) / 2 for i in range(3)], [p2[i] - p1[i] for i in range(3)]


def _faces(centre, side, pairs):
    """Axis-aligned rectangular faces: for (axis k, sign) the face at centre_k + sign*side_k/2, spanned by
    s_1, s_2 in [-1, 1] along the two other axes (in xyz order)."""
    s = symbols("s_1 s_2")
    ranges = {s[0]: (-1, 1), s[1]: (-1, 1)}
    out = []
    for k, sgn in pairs:
        others = [i for i in range(3) if i != k]
        fn = {"xyz"[k]: centre[k] + sgn * 0.5 * side[k]}
        for par, i in zip(s, others):
            fn["xyz"[i]] = centre[i] + 0.5 * par * side[i]
        fn = {a: fn[a] for a in "xyz"}
        fn.update(_normals("xyz", **{"xyz"[k]: sgn}))
        out.append(Edge(fn, ranges, side[others[0]] * side[others[1]]))
    return out


class Plane(Geometry3D):
    """Rectangle in the plane x = const; `normal` = +-1."""

    def __init__(self, point_1, point_2, normal):
        assert point_1[0] == point_2[0], "Points must have the same x coordinate"
        centre, side = _centre_sides(point_1, point_2)
        s = symbols("s_1 s_2")
        self.edges = [Edge({"x": centre[0], "y": centre[1] + 0.5 * s[0] * side[1], "z": centre[2] + 0.5 * s[1] * side[2],
                            "normal_x": 1e-10 + normal, "normal_y": 0, "normal_z": 0},
                           {s[0]: (-1, 1), s[1]: (-1, 1)}, side[1] * side[2])]
        self.sdf = normal * (centre[0] - _XYZ[0])
        self.bounds = _box_bounds(point_1, point_2)


class Tube3D(Geometry3D):
    """Channel along x: the four walls y = const, z = const."""

    def __init__(self, point_1, point_2):
        centre, side = _centre_sides(point_1, point_2)
        self.edges = _faces(centre, side, [(2, 1), (2, -1), (1, 1), (1, -1)])
        self.sdf = _slab_sdf(_XYZ[1:], centre[1:], [0.5 * side[1], 0.5 * side[2]])
        self.bounds = _box_bounds(point_1, point_2)


class Tube(Tube3D):
    pass


class CircularTube(Geometry3D):
    def __init__(self, center, radius, height):
        x, y, z = _XYZ
        h, theta = symbols("h theta")
        self.edges = [Edge({"x": center[0] + radius * cos(theta), "y": center[1] + radius * sin(theta),
                            "z": center[2] + 0.5 * h * height, "normal_x": 1 * cos(theta), "normal_y": 1 * sin(theta),
                            "normal_z": 0}, {h: (-1, 1), theta: (0, 2 * pi)}, height * 2 * pi * radius)]
        self.sdf = radius - sqrt((x - center[0]) ** 2 + (y - center[1]) ** 2)
        self.bounds = {"x": (center[0] - radius, center[0] + radius), "y": (center[1] - radius, center[1] + radius),
                       "z": (center[2] - height / 2, center[2] + height / 2)}


class Box(Geometry3D):
    def __init__(self, point_1, point_2):
        centre, side = _centre_sides(point_1, point_2)
        self.edges = _faces(centre, side, [(2, 1), (2, -1), (1, 1), (1, -1), (0, 1), (0, -1)])
        self.sdf = _slab_sdf(_XYZ, centre, [0.5 * s for s in side])
        self.bounds = _box_bounds(point_1, point_2)
        if _numeric(*point_1, *point_2):
            self._tree = ("prim", "box", tuple(point_1) + tuple(point_2))


class Sphere(Geometry3D):
    def __init__(self, center, radius):
        x, y, z = _XYZ
        v1, v2 = symbols("v_1 v_2")
        theta, phi = 2 * pi * v1, 2 * asin(sqrt(v2))      # uniform on the sphere
        r = [cos(theta) * sin(phi), sin(theta) * sin(phi), cos(phi)]
        norm = sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2)
        fn = {a: center[i] + radius * r[i] / norm for i, a in enumerate("xyz")}
        fn.update({"normal_" + a: r[i] / norm for i, a in enumerate("xyz")})
        self.edges = [Edge(fn, {v1: (0, 1), v2: (0, 1)}, 4 * pi * radius ** 2)]
        self.sdf = radius - sqrt((x - center[0]) ** 2 + (y - center[1]) ** 2 + (z - center[2]) ** 2)
        self.bounds = {a: (center[i] - radius, center[i] + radius) for i, a in enumerate("xyz")}
        if _numeric(*center, radius):
            self._tree = ("prim", "sphere", tuple(center) + (radius,))


class Cylinder(Geometry3D):
    def __init__(self, center, radius, height):
        x, y, z = _XYZ
        h, r, theta = symbols("h r theta")
        ranges = {h: (-1, 1), r: (0, 1), theta: (0, 2 * pi)}
        mantle = Edge({"x": center[0] + radius * cos(theta), "y": center[1] + radius * sin(theta),
                       "z": center[2] + 0.5 * h * height, "normal_x": 1 * cos(theta), "normal_y": 1 * sin(theta),
                       "normal_z": 0}, ranges, height * 2 * pi * radius)
        caps = [Edge({"x": center[0] + sqrt(r) * radius * cos(theta), "y": center[1] + sqrt(r) * radius * sin(theta),
                      "z": center[2] + sgn * 0.5 * height, **_normals("xyz", z=sgn)}, ranges, pi * radius ** 2)
                for sgn in (1, -1)]
        self.edges = [mantle] + caps
        rd = sqrt((x - center[0]) ** 2 + (y - center[1]) ** 2)
        zd = Abs(z - center[2])
        outside = sqrt(Min(0, radius - rd) ** 2 + Min(0, 0.5 * height - zd) ** 2)
        inside = -1 * Min(Abs(Min(0, rd - radius)), Abs(Min(0, zd - 0.5 * height)))
        self.sdf = -(outside + inside)
        self.bounds = {"x": (center[0] - radius, center[0] + radius), "y": (center[1] - radius, center[1] + radius),
                       "z": (center[2] - height / 2, center[2] + height / 2)}
