"""`Variables`: dict of named [N,1] tensors moving through the node graph, the
generic-path differentiation rule and the loss reductions (mirror of
idrlnet/variable.py; csv/vtu export is out of scope)."""
import enum
import itertools
import os
import pathlib
from collections import defaultdict
from typing import Dict, List, Union

import numpy as np
import torch

from .header import DIFF_SYMBOL

__all__ = ["Loss", "Variables", "DomainVariables", "export_var"]


class Loss(enum.Enum):
    L1 = "L1"
    square = "square"
    Identity = "Identity"


_LOSS_FN = {"L1": torch.abs, "square": lambda v: v ** 2, "Identity": lambda v: v}


def default_device():
    from . import DEVICE
    return DEVICE


class Variables(dict):
    def __sub__(self, other: "Variables") -> "Variables":  # idrlnet/variable.py:84-91
        return Variables({k: (self[k] if k in self else 0) - (other[k] if k in other else 0) for k in {**self, **other}})

    def weighted_loss(self, name: str, loss_function: Union[Loss, str]) -> "Variables":
        """sum_keys sum_n f(val) * [lambda_key] * area  (idrlnet/variable.py:30-80)."""
        kind = loss_function.name if isinstance(loss_function, Loss) else loss_function
        if kind not in _LOSS_FN:
            raise NotImplementedError(f"loss function {loss_function} is not defined!")
        f = _LOSS_FN[kind]
        loss = 0.0
        for key, val in self.items():
            if key.startswith("lambda_") or key == "area":
                continue
            if "lambda_" + key in self:
                loss = loss + torch.sum(f(val) * self["lambda_" + key] * self["area"])
            else:
                loss = loss + torch.sum(f(val) * self["area"])
        return Variables({name: loss})

    def subset(self, subset_keys: List[str]) -> "Variables":
        return Variables({k: self[k] for k in subset_keys if k in self})

    def to_torch_tensor_(self) -> "Variables":
        """numpy -> tensors of torch's default dtype (FP32 unless the user switched it, exactly what the reference's
        `torch.Tensor(val)` gives) on the active device; everything except `area` and `lambda_*` becomes a
        requires_grad leaf (idrlnet/variable.py:105-113)."""
        dev = default_device()
        for key, val in self.items():
            if not isinstance(val, torch.Tensor):
                t = torch.as_tensor(np.asarray(val), dtype=torch.get_default_dtype()).to(dev)
                if not key.startswith("lambda_") and key != "area":
                    t.requires_grad_()
                self[key] = t
        return self

    def to_ndarray_(self) -> "Variables":
        for key, val in self.items():
            if isinstance(val, torch.Tensor):
                self[key] = val.detach().cpu().numpy()
        return self

    def to_ndarray(self) -> "Variables":
        return Variables({k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in self.items()})

    def merge_tensor(self) -> torch.Tensor:
        return torch.cat(list(self.values()), dim=-1)

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor, variable_names: List[str]):
        cols = torch.split(tensor, 1, dim=-1)
        assert len(variable_names) == len(cols)
        return cls(zip(variable_names, cols))

    def differentiate_one_step_(self, independent_var: "Variables", required_derivatives: List[str]):
        """One order of derivatives by autograd: one torch.autograd.grad per
        (dependent, independent) pair (idrlnet/variable.py:161-196).  Generic path
        only -- the fused kernels obtain the same quantities as forward jets."""
        required = [d for d in required_derivatives if d not in self]
        req = set(tuple(r.split(DIFF_SYMBOL)) for r in required)
        dep = set(tuple(k.split(DIFF_SYMBOL)) for k in self.keys())
        todo = defaultdict(set)
        for dv, rd in itertools.product(dep, req):
            if len(rd) > len(dv) and rd[: len(dv)] == dv and rd[: len(dv) + 1] not in dep:
                todo[rd[len(dv)]].add(DIFF_SYMBOL.join(dv))
        new = Variables()
        for key, deps in todo.items():
            for v in deps:
                g = None
                if self[v].requires_grad:
                    g = torch.autograd.grad(self[v], independent_var[key], grad_outputs=torch.ones_like(self[v]),
                                            retain_graph=True, create_graph=True, allow_unused=True)[0]
                if g is not None:
                    g.requires_grad_()
                else:
                    g = torch.zeros_like(self[v], requires_grad=True)
                new[DIFF_SYMBOL.join([v, key])] = g
        self.update(new)

    def differentiate_(self, independent_var: "Variables", required_derivatives: List[str]):
        n = -1
        while n != len(self):  # idrlnet/variable.py:198-208
            n = len(self)
            self.differentiate_one_step_(independent_var, required_derivatives)

    def save(self, path, formats=None):
        formats = ["np"] if formats is None else formats
        if "np" in formats:
            np.savez(path, **self.to_ndarray())
        if "csv" in formats:
            import pandas as pd
            arr = self.to_ndarray()
            pd.DataFrame(np.concatenate(list(arr.values()), axis=-1), columns=list(arr.keys())).to_csv(
                path if str(path).endswith(".csv") else str(path) + ".csv", index=False)


DomainVariables = Dict[str, Variables]


def export_var(domain_var: DomainVariables, path="./inference_domain/results", formats=None):
    path = pathlib.Path(path)
    path.mkdir(exist_ok=True, parents=True)
    for key in domain_var:
        domain_var[key].save(os.path.join(path, f"{key}"), ["np", "csv"] if formats is None else
                             [f for f in formats if f != "vtu"])
