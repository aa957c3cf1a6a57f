"""Predefined receivers (mirror of idrlnet/callbacks.py).  Observability hooks are
outside the hot path; they receive the same Signals and payload keys as in the
reference.  TensorBoard writing is opt-in via IDRLNET_B200_TENSORBOARD=1 so that a
training step never synchronises the device just to log a scalar."""
import os
from typing import Dict


from .receivers import Receiver, Signal
from .variable import Variables

__all__ = ["GradientReceiver", "SummaryReceiver", "HandleResultReceiver"]


class GradientReceiver(Receiver):
    """Logs the gradient norm before the optimiser step (idrlnet/callbacks.py:13-30)."""

    def receive_notify(self, solver, message: Dict):
        if Signal.TRAIN_PIPE_END in message and getattr(solver, "summary_receiver", None) is not None:
            total = 0.0
            for nn in solver.netnodes:
                for p in nn.net.parameters():
                    if p.grad is not None:
                        total += float(p.grad.detach().norm(2)) ** 2
            w = getattr(solver.summary_receiver, "writer", None)
            if w is not None:
                w.add_scalar("gradient/total_norm", total ** 0.5, solver.global_step)


class SummaryReceiver(Receiver):
    """TensorBoard scalars `loss_overview`, `loss_component/*`, `optimizer/lr_i`
    (idrlnet/callbacks.py:33-51)."""

    def __init__(self, log_dir=None):
        self.log_dir = log_dir
        self.writer = None
        if os.environ.get("IDRLNET_B200_TENSORBOARD", "0") == "1":
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir)

    def receive_notify(self, solver, message: Dict):
        if self.writer is None:
            return
        if Signal.AFTER_COMPUTE_LOSS in message:
            comps = message[Signal.AFTER_COMPUTE_LOSS]
            self.writer.add_scalars("loss_overview", {k: float(v) for k, v in comps.items()}, solver.global_step)
            for k, v in comps.items():
                self.writer.add_scalar(f"loss_component/{k}", float(v), solver.global_step)
        if Signal.TRAIN_PIPE_END in message:
            for i, opt in enumerate(solver.optimizers):
                self.writer.add_scalar(f"optimizer/lr_{i}", opt.param_groups[0]["lr"], solver.global_step)


class HandleResultReceiver(Receiver):
    """Dumps the training-domain samples at SOLVE_END (idrlnet/callbacks.py:54-87);
    opt-in via IDRLNET_B200_DUMP_RESULTS=1."""

    def __init__(self, result_dir):
        self.result_dir = result_dir

    def receive_notify(self, solver, message: Dict):
        if Signal.SOLVE_END in message and os.environ.get("IDRLNET_B200_DUMP_RESULTS", "0") == "1":
            samples = solver.sample_variables_from_domains()
            in_var, _, _ = solver.generate_in_out_dict(samples)
            pred = solver.forward_through_all_graph(in_var, solver.outvar_dict_index)
            os.makedirs(self.result_dir, exist_ok=True)
            for k in pred:
                Variables({**in_var[k], **pred[k]}).save(os.path.join(self.result_dir, k), ["np"])
