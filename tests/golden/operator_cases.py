"""Integral-operator cases shared by the golden generator (reference) and the parity test (idrlnet_b200)."""
import numpy as np
import sympy as sp
import torch


def _net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(1, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))


def int1d_volterra(sc):
    """examples/Volterra_IDE/volterra_ide.py:34-45: rhs(x) = int_0^x exp(s - x) f(s) ds, f a net."""
    x, s, fs = sp.Symbol("x"), sp.Symbol("s"), sp.Symbol("fs")
    netnode = sc.NetNode(inputs=("x",), outputs=("f",), net=_net(), name="net")
    node = sc.Int1DNode(expression=sp.exp(s - x) * fs, var=s, lb=0, ub=x, expression_name="rhs",
                        funs={"fs": {"eval": netnode, "input_map": {"x": "s"}, "output_map": {"f": "fs"}}}, degree=10)
    xs = torch.linspace(0.0, 5.0, 37).reshape(-1, 1)
    out = node.evaluate({"x": xs})
    return {"inputs": sorted(node.inputs), "outputs": sorted(node.outputs), "rhs": out["rhs"].detach().numpy()}


def int1d_plain(sc):
    """No nets inside: int_{-1}^{x^2} (s^2 + x) ds with degree 20."""
    x, s = sp.Symbol("x"), sp.Symbol("s")
    node = sc.Int1DNode(expression=s ** 2 + x, var=s, lb=-1, ub=x ** 2, expression_name="val")
    xs = torch.linspace(-1.0, 2.0, 11).reshape(-1, 1)
    return {"val": node.evaluate({"x": xs})["val"].detach().numpy()}


def icnode(sc):
    rng = np.random.default_rng(3)
    cols = {k: torch.tensor(rng.standard_normal((50, 1)), dtype=torch.float32)
            for k in ("dxdy", "u", "v", "normal_x", "normal_y")}
    cols["area"] = torch.full((50, 1), 0.02)
    scalar = sc.ICNode("dxdy", dim=2, time=False)
    flux = sc.ICNode(["u", "v"], dim=2, time=False, reduce_name="flux")
    return {"scalar_inputs": sorted(scalar.inputs), "scalar_outputs": sorted(scalar.outputs),
            "scalar": scalar.evaluate(cols)["integral_dxdy"].detach().numpy(),
            "flux_inputs": sorted(flux.inputs), "flux": flux.evaluate(cols)["integral_flux"].detach().numpy()}


CASES = {"int1d_volterra": int1d_volterra, "int1d_plain": int1d_plain, "icnode": icnode}
