#!/usr/bin/env python
"""Drives the UNMODIFIED reference Solver (baseline/_ref) with a `JetNetNode` backed by libpinn_b200.so on one golden
problem and prints JSON: the loss, its components, the raw parameter gradients' agreement with the golden vectors the
stock reference produced, and how many `torch.autograd.grad` calls the reference made (0 with the jet node; one per
derivative symbol without it).  Run in its own process: `idrlnet.use_gpu` switches torch's default tensor type.

    python integration/run_reference_with_jets.py [simple_poisson|burgers] [--stock]
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from baseline.ref_harness import load_reference, strict_step  # noqa: E402

CASES = {
    "simple_poisson": dict(inputs=("x", "y"), outputs=("T",), derivatives=("T__x", "T__y", "T__x__x", "T__y__y"), net="net1"),
    "burgers": dict(inputs=("x", "t"), outputs=("u",), derivatives=("u__x", "u__t", "u__x__x"), net="net1"),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "simple_poisson"
    stock = "--stock" in sys.argv
    sc = load_reference()
    sc.use_gpu(0)
    import idrlnet
    assert idrlnet.GPU_ENABLED, "needs a CUDA device"
    import sympy
    from tests import golden_util as gu
    from integration.jet_net import jet_net_node
    problem, flat = gu.load(name)
    case = CASES[name]
    spec = next(n for n in problem["nets"] if n["name"] == case["net"])
    # the reference's own MLP holding the golden weights
    torch.manual_seed(0)
    node0 = sc.get_net_node(inputs=case["inputs"], outputs=case["outputs"], name="tmp", arch=sc.Arch.mlp)
    mlp = node0.net
    lins = [l for k, l in mlp.layers.items() if not k.endswith("_activation")]
    with torch.no_grad():
        for lin, lay in zip(lins, spec["layers"]):
            lin.bias.copy_(torch.as_tensor(lay["b"]))
            lin.weight_g.copy_(torch.as_tensor(lay["g"]).reshape(lin.weight_g.shape))
            lin.weight_v.copy_(torch.as_tensor(lay["v"]))
    if stock:
        net = sc.NetNode(inputs=case["inputs"], outputs=case["outputs"], net=mlp, name=case["net"])
    else:
        net = jet_net_node(sc, case["inputs"], case["outputs"], case["derivatives"], mlp, name=case["net"])
    pdes = [sc.ExpressionNode(expression=sympy.sympify(e), name=k) for k, e in problem["exprs"].items()]
    doms = []
    for d in problem["domains"]:
        pts = {k: torch.as_tensor(np.asarray(v, np.float32)).cuda() for k, v in d["points"].items()}
        for k, v in pts.items():
            if k != "area" and not k.startswith("lambda_"):
                v.requires_grad_()
        n = len(next(iter(pts.values())))
        cons = {k: (torch.as_tensor(np.asarray(v, np.float32)).cuda() if np.ndim(v) else torch.full((n, 1), float(v)))
                for k, v in d["targets"].items()}

        def make(pts=pts, cons=cons, d=d):
            @sc.datanode(name=d["name"], loss_fn=d.get("loss", "square"))
            class Dom(sc.SampleDomain):
                def sampling(self, *args, **kwargs):
                    return pts, cons
            return Dom()
        doms.append(make())
    solver = sc.Solver(sample_domains=tuple(doms), netnodes=[net], pdes=pdes, max_iter=1, loading=False,
                       network_dir=tempfile.mkdtemp(prefix="jetnode_"), opt_config=dict(optimizer="SGD", lr=0.0))
    calls = {"n": 0}
    real = torch.autograd.grad

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    torch.autograd.grad = counting
    import idrlnet.variable as var
    var.torch.autograd.grad = counting
    loss = strict_step(solver)  # forward graph + compute_loss + backward + (lr = 0) step, all the reference's code
    torch.autograd.grad = real
    out = {"case": name, "stock": stock, "loss": float(loss), "golden_loss32": float(flat["ref32/loss"]),
           "golden_loss64": float(flat["ref64/loss"]), "autograd_grad_calls": calls["n"],
           "components": {k: float(v) for k, v in solver.loss_component.items()}, "grad_err": {}}
    params = dict(mlp.named_parameters())
    for pname, g in gu.ref_grads(flat, "ref64").items():
        rest = pname.split("/", 1)[1]
        p = params[rest]
        got = p.grad.detach().cpu().numpy().reshape(g.shape)
        out["grad_err"][rest] = float(np.max(np.abs(got - g)) / max(np.max(np.abs(g)), 1e-30))
    native = [l.split()[-1] for l in open("/proc/self/maps") if "libpinn_b200.so" in l]
    out["native_so_loaded"] = sorted(set(native))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
