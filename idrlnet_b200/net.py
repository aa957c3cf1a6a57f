"""NetNode (mirror of idrlnet/net.py): a torch module labelled with input and
output symbol names."""
from contextlib import nullcontext
from typing import Dict, List, Tuple, Union

import torch

from .node import Node

__all__ = ["NetNode"]


class WrapEvaluate:
    """dict of [N,1] columns -> cat in dict order -> net -> dict of output columns;
    extra net columns are dropped (idrlnet/net.py:14-36)."""

    def __init__(self, binding_node: "NetNode"):
        self.binding_node = binding_node

    def __call__(self, inputs):
        as_dict = isinstance(inputs, dict)
        if as_dict:
            cols = [v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=torch.float32) for v in inputs.values()]
            inputs = torch.cat(cols, dim=1)
        with torch.no_grad() if self.binding_node.require_no_grad else nullcontext():
            out = self.binding_node.net(inputs)
        if as_dict:
            out = {key: out[:, i: i + 1] for i, key in enumerate(self.binding_node.outputs)}
        return out


class NetNode(Node):
    counter = 0

    def __init__(self, inputs: Union[Tuple, List[str]], outputs: Union[Tuple, List[str]], net: torch.nn.Module,
                 fixed: bool = False, require_no_grad: bool = False, is_reference=False, name=None, *args, **kwargs):
        self.is_reference = is_reference
        self.inputs = inputs
        self.outputs = outputs
        self.derivatives = []
        self.net = net
        self.require_no_grad = require_no_grad
        self.fixed = fixed
        if name is not None:
            self.name = name
        else:
            self.name = "net_{}".format(type(self).counter)
            type(self).counter += 1
        self.evaluate = WrapEvaluate(binding_node=self)

    def __str__(self):
        return super().__str__() + str(self.net)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        return self.net.load_state_dict(state_dict, strict)

    def state_dict(self, destination=None, prefix: str = "", keep_vars: bool = False):
        return self.net.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)
