"""Predefined equations (mirror of idrlnet/pde_op/equations.py).  Each class only builds SymPy
expressions and calls `make_nodes()`; all of them are polynomials in the jet symbols, i.e. fusable
by `compiler.py`."""
from sympy import Number

from ..pde import PdeNode
from ._common import _T, _X, _Y, _Z, _coords, _field

__all__ = ["DiffusionNode", "NavierStokesNode", "WaveNode", "BurgersNode", "SchrodingerNode", "AllenCahnNode"]


def _grad_sum(coef, f, variables):
    return sum((coef * f.diff(v)).diff(v) for v in variables)


class DiffusionNode(PdeNode):
    """T_t - div(D grad T) - Q  (idrlnet/pde_op/equations.py:28-47)."""

    def __init__(self, T="T", D="D", Q=0, dim=3, time=True, **kwargs):
        super().__init__(**kwargs)
        assert type(T) == str, "T should be string"
        self.T = T
        allv = [_X, _Y, _Z, _T]
        Tf, Df, Qf = _field(T, allv), _field(D, allv), _field(Q, allv)
        eq = -Qf + (Tf.diff(_T) if time else 0) - _grad_sum(Df, Tf, [_X, _Y, _Z][:dim])
        self.equations = {"diffusion_" + T: eq}
        self.make_nodes()


class NavierStokesNode(PdeNode):
    """Incompressible-form continuity + momentum (idrlnet/pde_op/equations.py:50-114)."""

    def __init__(self, nu=0.1, rho=1.0, dim=2.0, time=False, **kwargs):
        super().__init__(**kwargs)
        self.dim, self.time = dim, time
        assert dim in [2, 3], "dim should be 2 or 3"
        variables = _coords(dim, time)
        u, v, p = (_field(s, variables) for s in ("u", "v", "p"))
        w = _field("w", variables) if dim == 3 else Number(0)
        nu, rho = _field(nu, variables), _field(rho, variables)
        mu = rho * nu
        vel = [u, v, w]
        space = [_X, _Y, _Z]
        eqs = {"continuity": rho.diff(_T) + sum((rho * q).diff(s) for q, s in zip(vel, space))}
        for q, s, name in zip(vel, space, ("momentum_x", "momentum_y", "momentum_z")):
            if name == "momentum_z" and dim != 3:
                continue
            eqs[name] = ((rho * q).diff(_T) + sum(c * (rho * q).diff(r) for c, r in zip(vel, space))
                         + p.diff(s) - sum((mu * q.diff(r)).diff(r) for r in space))
        self.equations = eqs
        self.make_nodes()


class WaveNode(PdeNode):
    """u_tt - div(c^2 grad u)  (idrlnet/pde_op/equations.py:117-144)."""

    def __init__(self, u="u", c="c", dim=3, time=True, **kwargs):
        super().__init__(**kwargs)
        assert dim in [1, 2, 3], "dim should be 1, 2 or 3."
        assert type(u) == str, "u should be string"
        self.u, self.dim, self.time = u, dim, time
        variables = _coords(dim, time)
        uf, cf = _field(u, variables), _field(c, variables)
        self.equations = {"wave_equation": uf.diff(_T, 2) - _grad_sum(cf ** 2, uf, [_X, _Y, _Z])}
        self.make_nodes()


class BurgersNode(PdeNode):
    """u_t + u u_x - v u_xx  (idrlnet/pde_op/equations.py:147-160)."""

    def __init__(self, u: str = "u", v="v"):
        super().__init__()
        assert type(u) == str, "u needs to be string"
        variables = [_X, _T]
        uf, vf = _field(u, variables), _field(v, variables)
        self.equations = {f"burgers_{str(uf)}": uf.diff(_T) + uf * uf.diff(_X) - vf * uf.diff(_X).diff(_X)}
        self.make_nodes()


class SchrodingerNode(PdeNode):
    """idrlnet/pde_op/equations.py:163-179."""

    def __init__(self, u="u", v="v", c=0.5):
        super().__init__()
        assert type(u) == str and type(v) == str, "u, v should be string"
        self.c = c
        uf, vf = _field(u, [_X, _T]), _field(v, [_X, _T])
        self.equations = {
            "real": uf.diff(_T) + c * vf.diff(_X, 2) + (uf ** 2 + vf ** 2) * vf,
            "imaginary": vf.diff(_T) - c * uf.diff(_X, 2) - (uf ** 2 + vf ** 2) * uf,
        }
        self.make_nodes()


class AllenCahnNode(PdeNode):
    """u_t - gamma_1 u_xx - gamma_2 (u - u^3)  (idrlnet/pde_op/equations.py:182-197)."""

    def __init__(self, u="u", gamma_1=0.0001, gamma_2=5):
        super().__init__()
        assert type(u) == str, "u should be string"
        self.gama_1, self.gama_2 = gamma_1, gamma_2
        uf = _field(u, [_X, _T])
        self.equations = {"AllenCahn_" + str(uf): uf.diff(_T) - gamma_1 * uf.diff(_X, 2) - gamma_2 * (uf - uf ** 3)}
        self.make_nodes()
