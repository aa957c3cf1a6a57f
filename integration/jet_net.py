"""Reference-side binding: the CUDA jet kernels as a node of the UNMODIFIED reference.

The reference has no FFI; its seam is the node protocol (idrlnet/node.py:11-66).  A `NetNode` whose declared
outputs include derivative names (`T__x`, `T__x__x` ...) is chosen by the graph's longest-prefix rule
(idrlnet/graph.py:95-121) to serve those names, so `Variables.differentiate_` (idrlnet/variable.py:161-208) never
calls `torch.autograd.grad` for them (idrlnet/graph.py:172-175).  `jet_net_node` builds such a node around the
reference's own MLP module: its `net` is a `torch.nn.Module` that returns value AND derivative columns computed by
`pinn_jet_forward` of libpinn_b200.so (reverse pass: `pinn_jet_backward` + `pinn_step_finalize`), wrapped as one
autograd op w.r.t. the module's parameters (`idrlnet_b200.fused.JetFunction`).  Everything else -- `Solver`,
`PdeNode`s, loss, optimizer, checkpoints -- stays the reference's.

    import idrlnet.shortcut as sc                      # the stock reference
    from integration.jet_net import jet_net_node
    net = jet_net_node(sc, inputs=("x", "y"), outputs=("T",), derivatives=("T__x", "T__y", "T__x__x", "T__y__y"),
                       net=sc.get_net_node(inputs=("x", "y"), outputs=("T",), arch=sc.Arch.mlp, name="n").net, name="net1")
    sc.Solver(sample_domains=..., netnodes=[net], pdes=[...]).solve()
"""
from __future__ import annotations

from typing import Sequence

import torch

from idrlnet_b200 import compiler as cp
from idrlnet_b200 import fused


class JetModule(torch.nn.Module):
    """[N, d] coordinates -> [N, len(names)] columns: the value of every output, then the requested derivatives."""

    def __init__(self, mlp: torch.nn.Module, inputs: Sequence[str], outputs: Sequence[str], derivatives: Sequence[str],
                 precision: str = "fp32"):
        super().__init__()
        self.mlp = mlp  # registered: Solver hands `net.parameters()` to the optimizer and saves `state_dict()`
        self.inputs, self.base_outputs = tuple(inputs), tuple(outputs)
        self.names = tuple(outputs) + tuple(derivatives)
        tuples = []
        for name in derivatives:
            base, derivs = cp.split_name(name)
            if base not in outputs:
                raise ValueError(f"{name}: {base} is not an output of this net")
            tuples.append(derivs)
        self.plan = cp.plan_jets(list(inputs), tuples)
        self.columns = [(0, o) for o in range(len(outputs))]
        for name, t in zip(derivatives, tuples):
            self.columns.append((self.plan.channel(list(inputs), t), outputs.index(cp.split_name(name)[0])))
        self.precision = precision
        self._pack = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._pack is None:
            self._pack = fused.NetPack(self.mlp, precision=self.precision)  # raises if the net is not a supported MLP on CUDA
        pack = self._pack
        wb = pack.differentiable_effective()  # effective weights stay differentiable w.r.t. weight_g / weight_v (/ omega)
        cols = [x[:, i] for i in range(x.shape[1])]
        jets = fused.jet_mlp(self.plan, pack.act, cols, wb, pack.precision)  # [C * n_out, N]
        return torch.stack([jets[c * pack.n_out + o] for c, o in self.columns], dim=1)


def jet_net_node(sc, inputs, outputs, derivatives, net: torch.nn.Module, name: str, precision: str = "fp32", **kw):
    """A reference `NetNode` (idrlnet/net.py:39-99) serving `outputs` + `derivatives` from the CUDA jet kernel.
    `sc` is the reference's `idrlnet.shortcut` (or this package's: the node protocol is the same)."""
    module = JetModule(net, inputs, outputs, derivatives, precision)
    return sc.NetNode(inputs=tuple(inputs), outputs=module.names, net=module, name=name, **kw)
