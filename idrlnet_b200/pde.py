"""PdeNode / ExpressionNode (mirror of idrlnet/pde.py): named SymPy equations whose
derivative atoms become "u__x"-style symbols."""
from typing import Dict, List

from .header import DIFF_SYMBOL
from .node import Node
from .torch_util import _replace_derivatives
from .variable import Variables

__all__ = ["PdeNode", "ExpressionNode"]


class PdeEvaluate:
    def __init__(self, binding_pde):
        self.binding_pde = binding_pde

    def __call__(self, inputs) -> Variables:
        result = Variables()
        for node in self.binding_pde.sub_nodes:
            sub = {k: v for k, v in dict(inputs).items() if k in node.inputs or k in node.derivatives}
            result.update(node.evaluate(sub))
        return result


class PdeNode(Node):
    def __init__(self, suffix: str = "", **kwargs):
        self.suffix = "[" + suffix + "]" if len(suffix) > 0 else ""
        self.name = type(self).__name__ + self.suffix
        self.evaluate = PdeEvaluate(self)

    @property
    def equations(self) -> Dict:
        return self._equations

    @equations.setter
    def equations(self, equations: Dict):
        self._equations = equations

    def make_nodes(self) -> None:
        """One sub-node per equation; inputs/derivatives split by `__`
        (idrlnet/pde.py:65-79)."""
        self.sub_nodes: List[Node] = []
        free, names = set(), set()
        for name, eq in self.equations.items():
            torch_eq = _replace_derivatives(eq)
            symbols = [s.name for s in torch_eq.free_symbols]
            free.update(symbols)
            name = name + self.suffix
            self.sub_nodes.append(Node.new_node(name, torch_eq, symbols))
            names.add(name)
        self.inputs = [s for s in free if DIFF_SYMBOL not in s]
        self.derivatives = [s for s in free if DIFF_SYMBOL in s]
        self.outputs = list(names)

    def __str__(self):
        sub = "\n\n".join(str(n) + "Equation: \n" + str(self.equations.get(n.name, "")) for n in self.sub_nodes)
        return super().__str__() + "subnodes".center(30, "-") + "\n" + sub


class ExpressionNode(PdeNode):
    def __init__(self, expression, name, **kwargs):
        super().__init__(**kwargs)
        self.equations = {name: expression}
        self.name = name
        self.make_nodes()
