"""Pin the CPU oracle (oracle/pinn_oracle.py) against golden vectors produced by
executing the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from tests import golden_util as gu
from oracle import pinn_oracle as po

# param name in the reference's state dict -> (layer index, key in oracle grads)
_KEY = {"weight": "W", "weight_g": "g", "weight_v": "v", "bias": "b"}


def _check(name, dtype, tag, tol):
    problem, flat = gu.load(name)
    want = {d["name"]: [k for k in gu.ref_domain(flat, tag, d["name"])] for d in problem["domains"]}
    res = po.training_step(problem, dtype=dtype, want=want)
    # every variable the reference graph computed (outputs, derivative symbols, residuals)
    for d in problem["domains"]:
        ref = gu.ref_domain(flat, tag, d["name"])
        for k, v in ref.items():
            err = gu.normwise(res["domains"][d["name"]][k], v)
            assert err <= tol, f"{name}/{d['name']}/{k}: {err:.3e}"
    # per-domain and total loss
    for k, v in gu.ref_components(flat, tag).items():
        dom = k[: -len("_loss")]
        assert abs(res["loss_components"][dom] - v) <= tol * max(abs(v), 1e-30) * 10, k
    ref_loss = float(flat[f"{tag}/loss"])
    assert abs(res["loss"] - ref_loss) <= 10 * tol * abs(ref_loss)
    # parameter gradients (raw parameters, i.e. weight_g / weight_v / bias under weight-norm)
    for pname, g in gu.ref_grads(flat, tag).items():
        net, rest = pname.split("/", 1)
        parts = rest.split(".")  # layers.mlp_<i>.<param>
        li = int(parts[1].split("_")[1])
        og = res["grads"][net][li][_KEY[parts[2]]]
        assert og is not None, pname
        err = gu.normwise(og.reshape(g.shape), g)
        assert err <= tol * 20, f"{name}/{pname}: {err:.3e}"
    # parameters the reference leaves without gradient stay without gradient
    import json
    for pname in json.loads(str(flat[f"{tag}/grad_none"])):
        net, rest = pname.split("/", 1)
        parts = rest.split(".")
        li = int(parts[1].split("_")[1])
        assert res["grads"][net][li][_KEY[parts[2]]] is None, pname


@pytest.mark.parametrize("name", gu.CONFIGS)
def test_oracle_matches_reference_fp32(name):
    _check(name, torch.float32, "ref32", 2e-6)


@pytest.mark.parametrize("name", gu.CONFIGS)
def test_oracle_matches_reference_fp64(name):
    _check(name, torch.float64, "ref64", 1e-12)


def test_oracle_matches_reference_at_the_true_mlp_xl_shape():
    """Arch.mlp_xl 3-512x6-3 SiLU + weight-norm (idrlnet/architecture/mlp.py:235-246): every graph variable of the
    reference's FP32 and FP64 runs, the losses, and the FP64 run's parameter gradients (stored as float32, hence the
    2e-7 floor on top of the FP64 agreement)."""
    name = gu.XL
    problem, flat = gu.load(name)
    for dtype, tag, tol in ((torch.float32, "ref32", 2e-5), (torch.float64, "ref64", 1e-11)):
        want = {d["name"]: [k for k in gu.ref_domain(flat, tag, d["name"])] for d in problem["domains"]}
        res = po.training_step(problem, dtype=dtype, want=want)
        for d in problem["domains"]:
            for k, v in gu.ref_domain(flat, tag, d["name"]).items():
                err = gu.normwise(res["domains"][d["name"]][k], v)
                assert err <= tol, f"{d['name']}/{k}: {err:.3e}"
        ref_loss = float(flat[f"{tag}/loss"])
        assert abs(res["loss"] - ref_loss) <= 10 * tol * abs(ref_loss)
    n_checked = 0
    for pname, g in gu.ref_grads(flat, "ref64").items():
        net, rest = pname.split("/", 1)
        parts = rest.split(".")
        og = res["grads"][net][int(parts[1].split("_")[1])][_KEY[parts[2]]]
        assert gu.normwise(og.reshape(g.shape), g) <= 2e-7, pname
        n_checked += g.size
    assert n_checked == 1319942  # P of Arch.mlp_xl on (x, y, t) -> (u, v, p)
