"""SymPy expression -> postfix program of `pinn_sample_program` (include/pinn_b200.h, csrc/geometry.cu).

The reference evaluates an `Edge`'s coordinate / normal / measure expressions with a fresh NumPy `lambdify` on every
sampling call (idrlnet/geo_utils/geo.py:80-108).  Here each expression is compiled ONCE into a stack program over the
point's Philox-drawn parameters; one kernel launch then draws the parameters and writes every output column."""
from __future__ import annotations

from typing import Dict, List, Tuple

import sympy

from .. import native as nv


class Unsupported(Exception):
    pass


Op = Tuple[int, float]
_UNARY = {sympy.sin: "sin", sympy.cos: "cos", sympy.tan: "tan", sympy.exp: "exp", sympy.log: "log", sympy.tanh: "tanh",
          sympy.Abs: "abs", sympy.sign: "sign", sympy.Heaviside: "heaviside", sympy.asin: "asin", sympy.acos: "acos",
          sympy.atan: "atan", sympy.sinh: "sinh", sympy.cosh: "cosh"}


def compile_expr(expr, var_index: Dict[str, int], consts: Dict[str, float]) -> List[Op]:
    """Postfix program of `expr` over the variables `var_index` (name -> parameter slot) with the symbols in `consts`
    replaced by numbers.  Raises `Unsupported` for anything the device evaluator has no op for."""
    ops: List[Op] = []
    EX = nv.EX

    def emit(name: str, imm: float = 0.0):
        ops.append((EX[name], float(imm)))

    def rec(e):
        if isinstance(e, (int, float)):
            emit("const", e)
            return
        e = sympy.sympify(e)
        if e.is_Number or (e.is_number and not e.free_symbols):
            emit("const", float(e))
        elif e.is_Symbol:
            if e.name in var_index:
                emit("var", var_index[e.name])
            elif e.name in consts:
                emit("const", consts[e.name])
            else:
                raise Unsupported(f"free symbol {e.name}")
        elif e.is_Add or e.is_Mul:
            name, imm_name = ("add", "offset") if e.is_Add else ("mul", "scale")
            nums = [a for a in e.args if a.is_number and not a.free_symbols]
            rest = [a for a in e.args if not (a.is_number and not a.free_symbols)]
            if not rest:
                emit("const", float(e))
                return
            rec(rest[0])
            for a in rest[1:]:
                rec(a)
                emit(name)
            if nums:
                c = float(sympy.Add(*nums)) if e.is_Add else float(sympy.Mul(*nums))
                if e.is_Mul and c == -1.0:
                    emit("neg")
                else:
                    emit(imm_name, c)
        elif e.is_Pow:
            base, ex = e.args
            if ex.is_number and not ex.free_symbols:
                x = float(ex)
                rec(base)
                if x == 2.0:
                    emit("square")
                elif x == -1.0:
                    emit("rcp")
                elif x == 0.5:
                    emit("sqrt")
                elif x == -0.5:
                    emit("sqrt")
                    emit("rcp")
                elif x == 1.0:
                    pass
                else:
                    emit("const", x)
                    emit("pow")
            else:
                rec(base)
                rec(ex)
                emit("pow")
        elif isinstance(e, (sympy.Min, sympy.Max)):
            name = "min" if isinstance(e, sympy.Min) else "max"
            rec(e.args[0])
            for a in e.args[1:]:
                rec(a)
                emit(name)
        elif e.func in _UNARY and (len(e.args) == 1 or (e.func == sympy.Heaviside and e.args[1] == sympy.Rational(1, 2))):
            rec(e.args[0])
            emit(_UNARY[e.func])
        elif e.func == sympy.atan2:
            rec(e.args[0])
            rec(e.args[1])
            emit("atan2")
        else:
            raise Unsupported(f"no device op for {e.func}")

    rec(expr)
    # stack depth check (the library validates again)
    depth = peak = 0
    for op, _ in ops:
        if op in (EX["const"], EX["var"]):
            depth += 1
        elif op in (EX["add"], EX["sub"], EX["mul"], EX["div"], EX["min"], EX["max"], EX["pow"], EX["atan2"]):
            depth -= 1
        peak = max(peak, depth)
    if peak > nv.EXPR_STACK:
        raise Unsupported("expression too deep for the device stack")
    return ops


def pack_programs(programs: List[List[Op]]):
    """(ops array, begin array) for `pinn_sample_program`."""
    import ctypes as C
    total = sum(len(p) for p in programs)
    if total > nv.EXPR_MAX_OPS or len(programs) > nv.EXPR_MAX_OUT:
        raise Unsupported("programs too long for one launch")
    ops = (nv.PinnExprOp * max(total, 1))()
    begin = (C.c_int32 * (len(programs) + 1))()
    k = 0
    for i, prog in enumerate(programs):
        begin[i] = k
        for op, imm in prog:
            ops[k].op, ops[k].imm = op, imm
            k += 1
    begin[len(programs)] = k
    return ops, begin


def evaluate_host(prog: List[Op], var: List[float]) -> float:
    """Reference interpreter (tests): the same stack machine in Python floats."""
    import math
    EX = {v: k for k, v in nv.EX.items()}
    st: List[float] = []
    for op, imm in prog:
        n = EX[op]
        if n == "const":
            st.append(imm)
        elif n == "var":
            st.append(var[int(imm)])
        elif n in ("add", "sub", "mul", "div", "min", "max", "pow", "atan2"):
            b = st.pop()
            a = st.pop()
            fn = {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b, "div": lambda: a / b, "min": lambda: min(a, b),
                  "max": lambda: max(a, b), "pow": lambda: a ** b, "atan2": lambda: math.atan2(a, b)}[n]
            st.append(fn())
        else:
            a = st.pop()
            fn = {"neg": lambda: -a, "abs": lambda: abs(a), "sqrt": lambda: math.sqrt(a), "sin": lambda: math.sin(a),
                  "cos": lambda: math.cos(a), "tan": lambda: math.tan(a), "exp": lambda: math.exp(a), "log": lambda: math.log(a),
                  "tanh": lambda: math.tanh(a), "sign": lambda: float((a > 0) - (a < 0)),
                  "heaviside": lambda: 1.0 if a > 0 else (0.0 if a < 0 else 0.5), "square": lambda: a * a, "rcp": lambda: 1.0 / a,
                  "scale": lambda: a * imm, "offset": lambda: a + imm, "asin": lambda: math.asin(max(-1.0, min(1.0, a))),
                  "acos": lambda: math.acos(max(-1.0, min(1.0, a))), "atan": lambda: math.atan(a), "sinh": lambda: math.sinh(a),
                  "cosh": lambda: math.cosh(a)}[n]
            st.append(fn())
    assert len(st) == 1
    return float(st[0])
