"""SURVEY.md 8 b1, executed: the unmodified reference `Solver` (baseline/_ref) runs a `JetNetNode`
(integration/jet_net.py) backed by libpinn_b200.so -- 0 `torch.autograd.grad` calls, golden-level loss and
parameter gradients.  Own process: `idrlnet.use_gpu` changes torch's default tensor type."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "integration", "run_reference_with_jets.py")


def _run(*args):
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "idrlnet")):
        pytest.skip("baseline/_ref (pip --target install of the reference) is not present")
    out = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("case,n_derivs", [("simple_poisson", 4), ("burgers", 3)])
def test_jet_node_inside_the_unmodified_reference_solver(case, n_derivs):
    res = _run(case)
    assert res["autograd_grad_calls"] == 0  # every derivative symbol came from the kernel (idrlnet/graph.py:172-175)
    assert res["native_so_loaded"], "libpinn_b200.so was not loaded by the reference process"
    assert abs(res["loss"] - res["golden_loss64"]) <= 1e-5 * abs(res["golden_loss64"])
    assert max(res["grad_err"].values()) <= 2e-4, res["grad_err"]
    stock = _run(case, "--stock")  # the same script with the reference's own NetNode: autograd path
    assert stock["autograd_grad_calls"] >= n_derivs
    assert abs(stock["loss"] - res["loss"]) <= 1e-5 * abs(res["loss"])
