"""Optimizer / LR-scheduler configuration by name (mirror of idrlnet/optim.py)."""
import abc
import inspect
import math
from typing import Dict

import torch

__all__ = ["get_available_class", "Optimizable"]


def get_available_class(module, class_name) -> Dict[str, type]:
    return {n: c for n, c in inspect.getmembers(module)
            if inspect.isclass(c) and issubclass(c, class_name) and c is not class_name}


class Optimizable(metaclass=abc.ABCMeta):
    OPTIMIZER_MAP = get_available_class(torch.optim, torch.optim.Optimizer)
    # the reference reflects on `_LRScheduler` (idrlnet/optim.py:40-43), which on
    # current torch no scheduler subclasses any more; use the public base class.
    SCHEDULE_MAP = get_available_class(torch.optim.lr_scheduler, torch.optim.lr_scheduler.LRScheduler)

    @property
    def optimizers(self):
        return self._optimizers

    @optimizers.setter
    def optimizers(self, optimizers):
        self._optimizers = optimizers

    @property
    def schedulers(self):
        return self._schedulers

    @schedulers.setter
    def schedulers(self, schedulers):
        self._schedulers = schedulers

    @abc.abstractmethod
    def configure_optimizers(self):
        raise NotImplementedError

    def parse_configure(self, **kwargs):
        self.parse_optimizer(**kwargs)
        self.parse_lr_schedule(**kwargs)
        self.configure_optimizers()

    def parse_optimizer(self, **kwargs):  # idrlnet/optim.py:70-73
        cfg = dict(optimizer="Adam", lr=1e-3)
        cfg.update(kwargs.get("opt_config", {}))
        self.optimizer_config = cfg

    def parse_lr_schedule(self, **kwargs):  # idrlnet/optim.py:75-80
        cfg = dict(scheduler="ExponentialLR", gamma=math.pow(0.95, 0.001), last_epoch=-1)
        cfg.update(kwargs.get("schedule_config", {}))
        self.schedule_config = cfg

    def __str__(self):
        opt = str(self.optimizer_config) if "optimizer_config" in self.__dict__ else "optimizer is empty..."
        sch = str(self.schedule_config) if "schedule_config" in self.__dict__ else "scheduler is empty..."
        return "\n".join([opt, sch])
