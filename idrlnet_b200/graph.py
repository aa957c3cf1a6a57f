"""Per-domain computation graph (mirror of idrlnet/graph.py, PNG drawing left out).

Wiring rule (idrlnet/graph.py:95-121): a required name such as ``u__x__x`` is
served by the node whose declared output is its longest ``__``-prefix; a sampled
key wins only with a strictly longer prefix; whatever is left over is produced by
differentiation.  Evaluation order = reverse discovery order (idrlnet/graph.py:153-159).

Generic-path acceleration: a NetNode whose net the CUDA library supports is
evaluated by `fused.jet_mlp`, which returns the value AND every derivative
symbol the downstream nodes need in one kernel -- the seam of SURVEY.md 8(b1)
(a node that outputs ``u__x`` itself triggers no autograd call).
"""
from __future__ import annotations

from copy import copy
from typing import Dict, List, Optional, Sequence, Tuple

import sympy
import torch

from . import compiler as cp
from .header import DIFF_SYMBOL, logger
from .net import NetNode
from .node import Node
from .pde import PdeNode
from .variable import Variables

__all__ = ["VertexTaskPipeline", "Vertex"]


def _split(name: str) -> Tuple[str, ...]:
    return tuple(name.split(DIFF_SYMBOL))


class Vertex(Node):
    """Terminal vertex "<name>" asking for one output (identity evaluate)."""

    def __init__(self, req_name: str):
        if DIFF_SYMBOL in req_name:  # idrlnet/graph.py:75-77
            self.derivatives, self.inputs = (req_name,), tuple()
        else:
            self.inputs, self.derivatives = [req_name], tuple()
        self.outputs = tuple()
        self.name = f"<{req_name}>"
        self.evaluate = lambda x: x


class VertexTaskPipeline:
    MAX_STACK_ALLOWED = 100000

    def __init__(self, nodes: Sequence[Node], invar: Variables, req_names: List[str]):
        self.nodes = list(nodes)
        self.req_names = list(req_names)
        sampled = list(invar.keys())
        self.computable = set(sampled)
        stack: List[Node] = []
        pending: Dict[str, List[Node]] = {}
        for r in req_names:
            v = Vertex(r)
            stack.append(v)
            pending.setdefault(r, []).append(v)
        logger.debug("Constructing computation graph...")
        while pending:
            if len(stack) >= self.MAX_STACK_ALLOWED:
                raise ValueError("computation graph does not terminate")
            nxt: Dict[str, List[Node]] = {}
            for req, consumers in pending.items():
                rq = _split(req)
                best_len, best = -1, None
                for node in self.nodes:
                    if any(node is c for c in consumers):
                        continue  # a node never serves itself
                    for out in node.outputs:
                        o = _split(out)
                        if len(o) <= len(rq) and rq[: len(o)] == o and len(o) > best_len:
                            best_len, best = len(o), node
                for key in sampled:
                    k = _split(key)
                    if len(k) <= len(rq) and rq[: len(k)] == k and len(k) > best_len:
                        best_len, best = len(k), None
                if best_len <= 0:
                    raise Exception("Can't be computed: " + req)
                if best is not None:
                    logger.debug(f"{[c.name for c in consumers]}.{req} <---- {best.name}")
                    stack.append(best)
                    for p in list(best.inputs) + list(best.derivatives):
                        nxt.setdefault(p, []).append(best)
                self.computable.add(req)
            pending = nxt
        order: List[Node] = []
        while stack:
            n = stack.pop()
            if not any(n is m for m in order):
                order.append(n)
                self.computable |= set(n.outputs)
        self.evaluation_order_list = order
        self._jet_plans: Dict[int, "_JetServe"] = {}
        self._plan_jet_nodes(sampled)

    # ---- jet-serving NetNodes (generic path on the GPU) -------------------
    def _plan_jet_nodes(self, sampled: Sequence[str]) -> None:
        exprs = {}
        for n in self.evaluation_order_list:
            if isinstance(n, PdeNode):
                for s in n.sub_nodes:
                    if hasattr(s.evaluate, "tf_eq"):  # integral nodes (IntEq) are not expressions of the jets
                        exprs[s.name] = s.evaluate.tf_eq
        wanted = set()
        for n in self.evaluation_order_list:
            wanted.update(n.derivatives)
        for n in self.evaluation_order_list:
            if isinstance(n, NetNode):
                try:
                    self._check_no_chain_through(n, wanted)
                    js = _JetServe.build(n, wanted, sampled, exprs)
                except cp.NotFusable as e:
                    logger.debug(f"net {n.name}: torch path ({e})")
                    js = None
                if js is not None:
                    self._jet_plans[id(n)] = js

    def _check_no_chain_through(self, net: NetNode, wanted) -> None:
        """A jet-served net returns columns with NO autograd edge to the coordinates (the input derivatives are the
        jet channels).  That is only sound when every derivative symbol that depends on the net is a derivative OF
        one of the net's own outputs.  A wanted `flux__x` where `flux` is an ExpressionNode / PdeNode output built
        from `u`, or `v__x` of a second net fed with `u`, needs the chain rule through this net
        (idrlnet/variable.py:161-208 differentiates whatever the graph produced): such a pipeline keeps the torch
        path for this net instead of silently getting zeros from `allow_unused`."""
        producer = {}
        for n in self.evaluation_order_list:
            for out in n.outputs:
                producer.setdefault(out, n)
        mine = set(net.outputs)

        def reaches(name: str, seen: set) -> bool:
            base = cp.split_name(name)[0]
            if base in mine:
                return True
            node = producer.get(name) or producer.get(base)
            if node is None or id(node) in seen:
                return False
            seen.add(id(node))
            return any(reaches(q, seen) for q in list(node.inputs) + list(node.derivatives))

        for w in wanted:
            base, derivs = cp.split_name(w)
            if not derivs or base in mine:
                continue
            if reaches(base, set()):
                raise cp.NotFusable(f"{w}: derivative of a downstream quantity needs the chain rule through net {net.name}")

    def operation_order(self, invar: Variables):
        for node in self.evaluation_order_list:
            if not set(node.derivatives).issubset(invar.keys()):
                invar.differentiate_(independent_var=invar, required_derivatives=node.derivatives)
            js = self._jet_plans.get(id(node))
            args = {**invar.subset(node.inputs), **invar.subset(node.derivatives)}
            if js is not None and js.usable(args):
                invar.update(js.evaluate(args))
            else:
                invar.update(node.evaluate(args))

    def forward_pipeline(self, invar: Variables, req_names: List[str] = None) -> Variables:
        if req_names is None or set(req_names).issubset(set(self.computable)):
            outvar = copy(invar)
            self.operation_order(outvar)
            return outvar.subset(self.req_names if req_names is None else req_names)
        logger.info("The existing graph fails. Construct a temporary graph...")
        return VertexTaskPipeline(self.nodes, invar, req_names).forward_pipeline(invar)

    def display(self, filename: str = None):
        """Drawing the graph (networkx + matplotlib in the reference) is cosmetic; not provided."""
        return None


class _JetServe:
    """How a NetNode serves `u`, `u__x`, `u__x__x`... from one jet kernel call."""

    def __init__(self, node: NetNode, pack, plan: cp.JetPlan, names: Dict[str, Tuple[int, int, float]]):
        self.node, self.pack, self.plan, self.names = node, pack, plan, names

    @staticmethod
    def build(node: NetNode, wanted, sampled, exprs) -> Optional["_JetServe"]:
        from . import fused
        if node.require_no_grad:
            raise cp.NotFusable("require_no_grad net")
        p = next(node.net.parameters(), None)
        if p is None or not p.is_cuda:
            raise cp.NotFusable("net is not on a CUDA device")
        pack = fused.NetPack(node.net)
        inputs = list(node.inputs)
        if pack.n_in != len(inputs):
            raise cp.NotFusable("input count mismatch")
        # derivative variable -> (net input it acts through, slope)
        through: Dict[str, Tuple[str, float]] = {}
        for q in inputs:
            if q in sampled:
                through[q] = (q, 1.0)
            elif q in exprs:
                e = sympy.sympify(exprs[q])
                fs = list(e.free_symbols)
                if len(fs) != 1 or fs[0].name not in sampled:
                    raise cp.NotFusable(f"input {q} is not an affine map of one sampled coordinate")
                slope = e.diff(fs[0])
                if slope.free_symbols:
                    raise cp.NotFusable(f"input {q} is not affine")
                through[fs[0].name] = (q, float(slope))
            else:
                raise cp.NotFusable(f"input {q} is neither sampled nor an expression")
        mine, tuples = [], []
        for w in wanted:
            base, derivs = cp.split_name(w)
            if base not in node.outputs:
                continue
            for v in derivs:
                if v not in through:
                    raise cp.NotFusable(f"{w}: derivative w.r.t. {v} does not reach this net through an affine map")
            mine.append(w)
            tuples.append(tuple(through[v][0] for v in derivs))
        plan = cp.plan_jets(inputs, tuples)
        if len(node.outputs) > pack.n_out:
            raise cp.NotFusable("more outputs named than the net has")
        from . import native as nv
        import ctypes as C
        jets = nv.make_jets(plan.first_in, plan.n_second, plan.n_mixed, plan.n_high)
        pack.refresh_effective()
        if not nv.lib().pinn_supported(C.byref(pack.mlp_struct()), C.byref(jets)):
            raise cp.NotFusable("no kernel for this width / jet plan")
        names: Dict[str, Tuple[int, int, float]] = {}
        for o, out in enumerate(node.outputs):
            names[out] = (0, o, 1.0)
        for w, t in zip(mine, tuples):
            base, derivs = cp.split_name(w)
            scale = 1.0
            for v in derivs:
                scale *= through[v][1]
            names[w] = (plan.channel(inputs, t), node.outputs.index(base), scale)
        return _JetServe(node, pack, plan, names)

    def usable(self, args) -> bool:
        return all(isinstance(args.get(k), torch.Tensor) and args[k].is_cuda and args[k].dtype == torch.float32
                   for k in self.node.inputs)

    def evaluate(self, args) -> Dict[str, torch.Tensor]:
        from . import fused
        pack = self.pack
        wb = pack.differentiable_effective()
        jets = fused.jet_mlp(self.plan, pack.act, [args[k] for k in self.node.inputs], wb, pack.precision)
        n_out = pack.n_out
        out = {}
        for name, (c, o, scale) in self.names.items():
            col = jets[c * n_out + o].unsqueeze(1)
            out[name] = col if scale == 1.0 else col * scale
        return out
