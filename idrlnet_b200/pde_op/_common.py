"""Helpers shared by equations.py and operator.py."""
from sympy import Function, Number, symbols

_X, _Y, _Z, _T = symbols("x y z t")


def _coords(dim, time):
    v = [_X, _Y, _Z][: int(dim)]
    return v + ([_T] if time else [])


def _field(s, variables):
    """str -> unknown function of the coordinates; number -> constant."""
    if isinstance(s, str):
        return Function(s)(*variables)
    if isinstance(s, (int, float)):
        return Number(s)
    return s
