"""Node protocol (mirror of idrlnet/node.py): anything with .inputs, .derivatives,
.outputs, .name and .evaluate(dict) -> dict can be wired into the per-domain graph."""
from typing import Callable, List

from .header import DIFF_SYMBOL
from .torch_util import torch_lambdify

__all__ = ["Node"]


class Node(object):
    @property
    def inputs(self) -> List[str]:
        return self.__dict__.setdefault("_inputs", tuple())

    @inputs.setter
    def inputs(self, inputs):
        self._inputs = inputs

    @property
    def outputs(self) -> List[str]:
        return self.__dict__.setdefault("_outputs", tuple())

    @outputs.setter
    def outputs(self, outputs):
        self._outputs = outputs

    @property
    def derivatives(self) -> List[str]:
        return self.__dict__.setdefault("_derivatives", [])

    @derivatives.setter
    def derivatives(self, derivatives):
        self._derivatives = derivatives

    @property
    def evaluate(self) -> Callable:
        return self._evaluate

    @evaluate.setter
    def evaluate(self, evaluate: Callable):
        self._evaluate = evaluate

    @property
    def name(self) -> str:
        return self.__dict__.setdefault("_name", "Node" + str(id(self)))

    @name.setter
    def name(self, name: str):
        self._name = name

    @classmethod
    def new_node(cls, name: str = None, tf_eq=None, free_symbols: List[str] = None, *args, **kwargs) -> "Node":
        """A node evaluating one SymPy expression (idrlnet/node.py:68-85)."""
        node = cls()
        node.evaluate = LambdaTorchFun(free_symbols, tf_eq, name)
        node.inputs = [x for x in free_symbols if DIFF_SYMBOL not in x]
        node.derivatives = [x for x in free_symbols if DIFF_SYMBOL in x]
        node.outputs = [name]
        node.name = name
        return node

    def __str__(self):
        return ("Basic properties:\n" f"name: {self.name}\n" f"inputs: {self.inputs}\n"
                f"derivatives: {self.derivatives}\n" f"outputs: {self.outputs}\n")


class LambdaTorchFun:
    """Lambdified expression called with keyword tensors (idrlnet/node.py:98-109)."""

    def __init__(self, free_symbols, tf_eq, name):
        self.lambda_tf_eq = torch_lambdify(free_symbols, tf_eq)
        self.tf_eq = tf_eq
        self.name = name
        self.free_symbols = free_symbols

    def __call__(self, var):
        return {self.name: self.lambda_tf_eq(**dict(var))}
