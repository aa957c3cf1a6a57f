"""The SymPy -> postfix compiler behind `pinn_sample_program` (idrlnet_b200/geo_utils/expr_prog.py): every edge
expression and composed signed distance of the shipped shapes, evaluated by the host interpreter of the SAME programs,
against NumPy lambdify of the original expressions (what idrlnet/geo_utils/geo.py:80-108 evaluates)."""
import numpy as np
import pytest
import sympy as sp

import idrlnet_b200.shortcut as sc
from idrlnet_b200.geo_utils import expr_prog as xp
from idrlnet_b200.geo_utils.sympy_np import lambdify_np
from idrlnet_b200 import native as nv

t = sp.Symbol("t")
SHAPES = {
    "line": lambda: sc.Line1D(-1.0, 1.0),
    "rect": lambda: sc.Rectangle((-1.0, -0.5), (1.0, 0.5)),
    "circle": lambda: sc.Circle((0.2, -0.1), 0.7),
    "channel": lambda: sc.Channel2D((-1.0, -0.5), (1.0, 0.5)) if hasattr(sc, "Channel2D") else sc.Rectangle((0.0, 0.0), (1.0, 1.0)),
    "csg": lambda: sc.Rectangle((-1.0, -1.0), (1.0, 1.0)) - sc.Circle((0.0, 0.0), 0.3),
    "box": lambda: sc.Box((-1.0, -1.0, -1.0), (1.0, 2.0, 1.5)) if hasattr(sc, "Box") else sc.Rectangle((0.0, 0.0), (1.0, 1.0)),
    "sphere": lambda: sc.Sphere((0.1, 0.2, 0.3), 0.8) if hasattr(sc, "Sphere") else sc.Circle((0.0, 0.0), 1.0),
}


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_edge_programs_match_numpy(name):
    geo = SHAPES[name]()
    entry = geo._edge_programs(50, {t: (0.0, 2.0)})
    if name == "sphere" and entry is None:
        pytest.skip("the sphere's six column programs exceed one launch: it takes the torch lowering")
    assert entry is not None, "every expression of this shape has a device op"
    keys, plan = entry
    rng = np.random.default_rng(0)
    for edge, (n, ops, begin, lo, hi, n_vars) in zip(geo.edges, plan):
        ranges = {**edge.ranges, t: (0.0, 2.0)}
        drawn = [(k if isinstance(k, str) else k.name, v) for k, v in ranges.items() if isinstance(v, tuple)]
        assert n_vars == len(drawn)
        for _ in range(8):
            var = [rng.uniform(v[0], v[1]) for _, v in drawn]
            cols = {k: np.array([[x]]) for (k, _), x in zip(drawn, var)}
            for k, v in ranges.items():
                if not isinstance(v, tuple):
                    cols[k if isinstance(k, str) else k.name] = np.array([[float(v)]])
            names = list(cols)
            want = {"area": float(lambdify_np(edge.area, names)(**cols)) / max(n, 1)}
            for k, fn in edge.functions.items():
                want[k] = float(np.asarray(lambdify_np(fn, names)(**cols)).ravel()[0])
            want["t"] = cols["t"][0, 0]
            pos = {a: np.array([[want[a]]]) for a in edge.axes}
            want["sdf"] = float(np.asarray(lambdify_np(geo.sdf, list(pos))(**pos)).ravel()[0])
            for o, k in enumerate(keys):
                prog = [(ops[i].op, ops[i].imm) for i in range(begin[o], begin[o + 1])]
                got = xp.evaluate_host(prog, var)
                assert abs(got - want[k]) <= 2e-6 * max(1.0, abs(want[k])), (name, k, got, want[k])
            assert abs(want["sdf"]) < 1e-6 or name == "csg"  # an edge sample lies on the solid's own boundary


def test_unsupported_expressions_fall_back():
    x = sp.Symbol("x")
    with pytest.raises(xp.Unsupported):
        xp.compile_expr(sp.Piecewise((x, x > 0), (0, True)), {"x": 0}, {})
    with pytest.raises(xp.Unsupported):
        xp.compile_expr(sp.Symbol("unknown") + x, {"x": 0}, {})
    long = sum(sp.sin(x + i) for i in range(nv.EXPR_MAX_OPS // 2))
    with pytest.raises(xp.Unsupported):
        xp.pack_programs([xp.compile_expr(long, {"x": 0}, {})])
