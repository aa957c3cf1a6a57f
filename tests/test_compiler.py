"""Jet planner and residual compiler (host logic)."""
import pytest
import sympy

from idrlnet_b200 import compiler as cp
from idrlnet_b200 import native as nv
import idrlnet_b200.shortcut as sc


def _exprs(*pdes):
    out = {}
    for p in pdes:
        for s in p.sub_nodes:
            out[s.name] = s.evaluate.tf_eq
    return out


def test_equation_strings_match_the_reference():
    """SURVEY.md A.1 lists the reference's lambdified sources."""
    ns = _exprs(sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=True))
    assert sympy.simplify(ns["momentum_x"] - sympy.sympify(
        "p__x + 1.0*u*u__x + 1.0*u__t - 0.01*u__x__x + 1.0*u__y*v - 0.01*u__y__y")) == 0
    ac = _exprs(sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5))
    assert sympy.simplify(ac["AllenCahn_u"] - sympy.sympify("5*u**3 - 5*u + u__t - 0.0001*u__x__x")) == 0
    assert str(_exprs(sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False))["diffusion_T"]) == "-1.0*T__x__x - 1.0*T__y__y"
    assert "burgers_u" in _exprs(sc.BurgersNode(u="u", v=0.01))
    assert "normal_gradient_T" in _exprs(sc.NormalGradient("T", dim=2, time=False))


def test_plan_and_symbol_ids_burgers():
    ex = _exprs(sc.BurgersNode(u="u", v=0.01))
    cd = cp.compile_domain(["burgers_u"], ex, [cp.NetSpec("n", ("x", "t"), ("u",))], ["x", "t", "area"])
    assert cd.plan.first_in == (0, 1) and cd.plan.n_second == 1 and cd.plan.n_channels == 4
    terms = sorted(cd.equations[0].terms, key=lambda t: t[1])
    ids = [t[1] for t in terms]
    # u*u__x -> [0, 4]; u__t -> [8]; u__x__x -> [12]
    assert [0, 4] in ids and [8] in ids and [12] in ids


def test_plan_orders_second_order_inputs_first():
    cd = cp.compile_domain(["eq"], {"eq": "u__t + u__y__y"}, [cp.NetSpec("n", ("x", "y", "t"), ("u",))], ["x", "y", "t"])
    assert cd.plan.first_in == (1, 2) and cd.plan.n_second == 1


def test_ldc_three_equations_with_aux_lambda():
    ex = _exprs(sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=False))
    cd = cp.compile_domain(["continuity", "momentum_x", "momentum_y"], ex, [cp.NetSpec("n", ("x", "y"), ("u", "v", "p"))],
                           ["x", "y", "sdf", "area"])
    assert cd.plan.first_in == (0, 1) and cd.plan.n_second == 2 and len(cd.equations) == 3
    assert cd.n_terms == 2 + 5 + 5 and not cd.aux_keys


def test_normal_gradient_uses_aux_symbols():
    ex = _exprs(sc.NormalGradient("T", dim=2, time=False))
    cd = cp.compile_domain(["normal_gradient_T"], ex, [cp.NetSpec("n", ("x", "y"), ("T",))],
                           ["x", "y", "normal_x", "normal_y", "area"])
    assert sorted(cd.aux_keys) == ["normal_x", "normal_y"]
    for _, ids in cd.equations[0].terms:
        assert any(i >= nv.SYM_AUX0 for i in ids)


def test_expression_nodes_are_substituted():
    x = sympy.Symbol("x")
    u = sympy.Function("u")(x)
    du = sc.ExpressionNode(expression=u.diff(x), name="diff_u")
    cd = cp.compile_domain(["res"], {**_exprs(du), "res": "diff_u - 2*u"}, [cp.NetSpec("n", ("x",), ("u",))], ["x"])
    assert sorted(i for _, ids in cd.equations[0].terms for i in ids) == [0, 4]


@pytest.mark.parametrize("targets,exprs,nets,sampled,why", [
    (["e"], {"e": "sin(u)"}, [("n", ("x",), ("u",))], ["x"], "polynomial"),
    (["e"], {"e": "u__x__y + u__y__t"}, [("n", ("x", "y", "t"), ("u",))], ["x", "y", "t"], "more than one mixed"),
    (["e"], {"e": "u__x__x__y"}, [("n", ("x", "y"), ("u",))], ["x", "y"], "mixed, order > 2"),
    (["e"], {"e": "u__x__x__x__x__x"}, [("n", ("x",), ("u",))], ["x"], "order > 4"),
    (["e"], {"e": "u__x__x__x + u__y__y__y"}, [("n", ("x", "y"), ("u",))], ["x", "y"], "more than one input"),
    (["e"], {"e": "u__x__x__x + u__x__y"}, [("n", ("x", "y"), ("u",))], ["x", "y"], "mixed and third"),
    (["e"], {"e": "u - up"}, [("a", ("x",), ("u",)), ("b", ("xp",), ("up",))], ["x", "xp"], "nets"),
    (["e"], {"e": "u", "xp": "x + 2"}, [("n", ("xp",), ("u",))], ["x"], "upstream"),
    (["e"], {"e": "u*q"}, [("n", ("x",), ("u",))], ["x"], "cannot"),
])
def test_not_fusable(targets, exprs, nets, sampled, why):
    with pytest.raises(cp.NotFusable) as e:
        cp.compile_domain(targets, exprs, [cp.NetSpec(*n) for n in nets], sampled)
    assert why in str(e.value)


def test_mixed_and_high_order_channels_are_planned():
    """one mixed second derivative, or the pure third / fourth derivative along one input
    (examples/euler_beam/euler_beam.py:47-54), get their own jet channels."""
    cd = cp.compile_domain(["e"], {"e": "u__x__x + u__y__y + u__y__x"}, [cp.NetSpec("n", ("x", "y"), ("u",))], ["x", "y"])
    assert cd.plan.first_in == (0, 1) and cd.plan.n_second == 2 and cd.plan.n_mixed == 1 and cd.plan.n_channels == 6
    assert cd.plan.channel(("x", "y"), ("y", "x")) == cd.plan.channel(("x", "y"), ("x", "y")) == 5
    assert cd.plan.channel_names(("x", "y"))[5] == ("x", "y")
    x = sympy.Symbol("x")
    y = sympy.Function("y")(x)
    beam = sc.ExpressionNode(name="dddd_y", expression=y.diff(x).diff(x).diff(x).diff(x) + 1)
    cd = cp.compile_domain(["dddd_y"], _exprs(beam), [cp.NetSpec("n", ("x",), ("y",))], ["x"])
    assert cd.plan.first_in == (0,) and cd.plan.n_second == 1 and cd.plan.n_high == 2 and cd.plan.n_channels == 5
    assert sorted(i for _, ids in cd.equations[0].terms for i in ids) == [4 * nv.MAX_OUT]
    # the high-order input leads even when another input sorts before it
    cd = cp.compile_domain(["e"], {"e": "u__t__t__t + u__x__x"}, [cp.NetSpec("n", ("x", "t"), ("u",))], ["x", "t"])
    assert cd.plan.first_in == (1, 0) and cd.plan.n_second == 2 and cd.plan.n_high == 1
    assert cd.plan.channel(("x", "t"), ("t", "t", "t")) == 5


def test_domain_struct_round_trip():
    ex = _exprs(sc.AllenCahnNode())
    cd = cp.compile_domain(["AllenCahn_u"], ex, [cp.NetSpec("n", ("x", "t"), ("u",))], ["x", "t", "area"])
    d = cp.build_domain_struct(cd, 77, lambda k: {"x": 4096, "t": 8192}[k], {"AllenCahn_u": 0.0},
                               {"AllenCahn_u": cp.Ptr(12288)}, 0.5, "L1", 2.0)
    assert d.n_points == 77 and d.coords[0] == 4096 and d.coords[1] == 8192 and d.n_eq == 1
    assert d.eq[0].lambda_ == 12288 and d.eq[0].target is None and d.loss_kind == 1 and d.sigma == 2.0
    assert d.n_terms == 4 and d.eq[0].term_end == 4
    cubic = [t for t in d.terms[:4] if t.n_fac == 3][0]
    assert abs(cubic.coef - 5.0) < 1e-6 and list(cubic.fac[:3]) == [0, 0, 0]
