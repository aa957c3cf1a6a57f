"""Sampling cases shared by the golden generator (run against the reference) and the parity test
(run against idrlnet_b200): each case is `name -> callable(sc, sp)` returning a dict of arrays."""
import math


def _cases():
    def line1d(sc, sp):
        g = sc.Line1D(-1.0, 1.5)
        return {"b": g.sample_boundary(10), "i": g.sample_interior(40)}

    def line(sc, sp):
        return {"b": sc.Line((0.0, 0.0), (1.0, 2.0), normal=-1).sample_boundary(30)}

    def tube2d(sc, sp):
        g = sc.Tube2D((-1, -0.5), (1, 0.5))
        return {"b": g.sample_boundary(20), "i": g.sample_interior(30)}

    def rectangle(sc, sp):
        x, y = sp.symbols("x y")
        g = sc.Rectangle((-1.0, -2.0), (1.0, 2.0))
        return {"b": g.sample_boundary(25, sieve=((y > -2.0) & (y < 2.0))), "i": g.sample_interior(20, sieve=(x > 0)),
                "b_eq": g.sample_boundary(25, sieve=sp.Eq(y, 2.0))}

    def circle(sc, sp):
        g = sc.Circle((0.2, -0.1), 0.7)
        return {"b": g.sample_boundary(40), "i": g.sample_interior(60)}

    def heart(sc, sp):
        g = sc.Heart((0.0, 0.5), 0.5)
        return {"b": g.sample_boundary(40), "i": g.sample_interior(100)}

    def triangle(sc, sp):
        g = sc.Triangle((0.0, 0.0), (2.0, 0.3), (0.5, 1.5))
        return {"b": g.sample_boundary(30), "i": g.sample_interior(60)}

    def polygon(sc, sp):
        g = sc.Polygon([(0, 0), (2, 0), (2, 1), (1, 2), (0, 1)])
        return {"b": g.sample_boundary(20), "i": g.sample_interior(50)}

    def plane(sc, sp):
        return {"b": sc.Plane((0.5, -1, -1), (0.5, 1, 2), normal=1).sample_boundary(10)}

    def tube3d(sc, sp):
        g = sc.Tube3D((-1, -0.5, -0.25), (1, 0.5, 0.25))
        return {"b": g.sample_boundary(10), "i": g.sample_interior(40)}

    def circular_tube(sc, sp):
        g = sc.CircularTube((0, 0, 0), 0.5, 2.0)
        return {"b": g.sample_boundary(10), "i": g.sample_interior(40)}

    def box(sc, sp):
        g = sc.Box((-1, -0.5, 0), (1, 0.5, 0.75))
        return {"b": g.sample_boundary(10), "i": g.sample_interior(40)}

    def sphere(sc, sp):
        g = sc.Sphere((0.1, 0.2, 0.3), 0.8)
        return {"b": g.sample_boundary(10), "i": g.sample_interior(40)}

    def cylinder(sc, sp):
        g = sc.Cylinder((0, 0, 0), 0.5, 1.5)
        return {"b": g.sample_boundary(10), "i": g.sample_interior(60)}

    def csg(sc, sp):
        r, c = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Circle((0.3, 0.0), 0.5)
        c2 = sc.Circle((-0.8, 0.9), 0.6)
        return {"sub_b": (r - c).sample_boundary(30), "sub_i": (r - c).sample_interior(50),
                "add_b": (r + c2).sample_boundary(30), "add_i": (c + c2).sample_interior(50),
                "and_b": (r & c2).sample_boundary(30), "and_i": (r & c2).sample_interior(50),
                "poly_i": (sc.Polygon([(0, 0), (2, 0), (1, 2)]) - c).sample_interior(30)}

    def csg_invert(sc, sp):  # the reference's __invert__ raises (bounds is None, idrlnet/geo_utils/geo.py:282-288)
        return {"inv_b": (~sc.Circle((0.3, 0.0), 0.5)).sample_boundary(30)}

    def csg_triangle(sc, sp):
        r, t = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Triangle((0.0, 0.0), (1.5, 0.0), (0.0, 1.5))
        return {"and_b": (r & t).sample_boundary(30), "and_i": (r & t).sample_interior(50)}

    def holes(sc, sp):  # parameterised geometry (examples/holes_heat_flux/holes_heat_flux.py:10-37)
        r1, r2 = sp.symbols("r1 r2")
        plate = sc.Tube2D((-1, -0.5), (1, 0.5))
        geo = plate - sc.Circle(center=(-0.6, 0), radius=r1) - sc.Circle(center=(0.0, 0.0), radius=r2)
        pr = {r1: (0.05, 0.2), r2: 0.1}
        return {"b": geo.sample_boundary(40, param_ranges=pr), "i": geo.sample_interior(60, param_ranges=pr),
                "in": sc.Line((-1, -0.5), (-1, 0.5), normal=1).sample_boundary(20, param_ranges=pr)}

    def moved(sc, sp):
        g = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)).translation((0.5, -0.25)).scaling(1.5)
        c = sc.Circle((0.0, 0.0), 0.5).scaling(2.0, center=(1.0, 1.0))
        return {"b": g.sample_boundary(20), "i": g.sample_interior(20), "cb": c.sample_boundary(20),
                "ci": c.sample_interior(20)}

    def rotated_interior(sc, sp):  # interior only: the reference's Edge.rotation is not a rotation
        g = sc.Rectangle((-1.0, -0.5), (1.0, 0.5)).rotation(math.pi / 6)
        return {"i": g.sample_interior(60)}

    def low_discrepancy(sc, sp):
        g = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))
        return {"i": g.sample_interior(40, low_discrepancy=True), "b": g.sample_boundary(15, low_discrepancy=True),
                "box": sc.Box((0, 0, 0), (1, 1, 1)).sample_interior(50, low_discrepancy=True)}

    return {k: v for k, v in locals().items() if callable(v)}


CASES = _cases()
SEED = 20240
