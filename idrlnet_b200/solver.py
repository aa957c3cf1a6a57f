"""Solver (mirror of idrlnet/solver.py) with the fused CUDA training step behind it.

Public surface kept: constructor arguments, `solve`, `train_pipe`, `compute_loss`,
`infer_step`, `forward_through_all_graph`, `generate_in_out_dict`,
`sample_variables_from_domains`, `set_domain_parameter` / `get_domain_parameter`,
`register_receiver`, `sample_domains` (setter rebuilds the graphs), `netnodes`,
`global_step`, `loss_component`, `save` / `load` / `init_load` (same `model.ckpt`
layout), the seven Signals.

What changes underneath: at graph-build time every domain whose loss terms compile
to polynomials over one supported net (`compiler.compile_domain`) is bound to the
fused kernels; a training step for those domains is `pinn_step_fused` +
`pinn_step_finalize` (forward jets, residual, loss and reverse pass in one kernel
per domain) instead of forward graph + `compute_loss` + `loss.backward()`
(idrlnet/solver.py:329-338).  `Solver(precision="tf32")` runs the layer GEMMs of wide nets on
the tensor cores in TF32 (stated tolerance 1e-2); the default "fp32x3" (tensor cores, 3xTF32 operand splitting) and "fp32"
(FP32 FMA) hold 1e-5.  Everything else runs the generic graph.  Under
`torch.distributed` the points of every domain are sharded over the ranks and
the flat gradient (+ loss scalars) is all-reduced once per step.

Reference semantics are the default: like idrlnet/solver.py:341-344, `train_pipe`
re-samples every domain after the optimiser step and returns / logs the loss of that
second, gradient-free evaluation (run here by the loss-only mode of the kernels,
`pinn_loss_fused`, about a third of a full step), so an unchanged example sees the
same returned losses, the same Signals (BEFORE/AFTER_COMPUTE_LOSS fire twice per
step) and the same consumption of the NumPy sampling stream as with the reference.
`Solver(post_step_loss=False)` (or IDRLNET_B200_POST_STEP_LOSS=0) is the opt-in fast
mode: one sampling round per step, the returned loss is the one the gradient was
taken of.
"""
from __future__ import annotations

import os
import pathlib
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch

from . import compiler as cp
from .data import DataNode, SampleDomain
from .graph import VertexTaskPipeline
from .header import logger
from .net import NetNode
from .optim import Optimizable
from .pde import PdeNode
from .receivers import Notifier, Receiver, Signal
from .variable import DomainVariables, Variables

__all__ = ["Solver"]


_NVTX = os.environ.get("IDRLNET_B200_NVTX", "0") not in ("0", "", "false", "False")


class _Range:
    """NVTX range around a phase of the step (IDRLNET_B200_NVTX=1; shows up in nsys / ncu --nvtx timelines):
    `sample`, `fused_step`, `generic_graph`, `all_reduce`, `optimizer`, `post_step_loss`."""
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX and torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX and torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


def _dist():
    d = torch.distributed
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        return d
    return None


class Solver(Notifier, Optimizable):
    def __init__(self, sample_domains: Tuple[Union[DataNode, SampleDomain], ...], netnodes: List[NetNode],
                 pdes: Optional[List] = None, network_dir: str = "./network_dir", summary_dir: Optional[str] = None,
                 max_iter: int = 1000, save_freq: int = 100, print_freq: int = 10, loading: bool = True,
                 init_network_dirs: Optional[List[str]] = None, opt_config: Dict = None, schedule_config: Dict = None,
                 result_dir="train_domain/results", fused: Optional[bool] = None, post_step_loss: Optional[bool] = None,
                 precision: Optional[str] = None, shard_points: bool = True, device_sampling: Optional[bool] = None,
                 sampling_seed: int = 0, **kwargs):
        from . import callbacks
        self.network_dir = network_dir
        self.domain_losses = {d.name: d.loss_fn for d in sample_domains}
        self.netnodes: List[NetNode] = netnodes
        self.fused_enabled = fused
        if post_step_loss is None:  # reference behaviour unless the fast mode is asked for
            post_step_loss = os.environ.get("IDRLNET_B200_POST_STEP_LOSS", "1") not in ("0", "false", "False")
        self.post_step_loss = bool(post_step_loss)
        # arithmetic of the layer GEMMs of wide (H > 32) nets: "fp32x3" (tensor cores, 3xTF32 splitting: FP32-grade, the
        # default), "fp32" (FP32 FMA), "tf32" / "tf32_streamed" (tensor cores, 1e-2); None -> IDRLNET_B200_PRECISION or "fp32x3"
        self.precision = precision
        # under torch.distributed every sampled domain is split [rank::world] (all ranks draw the same seeded points, as
        # an unchanged example does); shard_points=False is for callers whose samplers already return this rank's points
        self.shard_points = shard_points
        # draw the points of every `geo.sample_interior` / `geo.sample_boundary` call on the GPU (Philox
        # streams, rank-offset under torch.distributed: each rank draws ITS OWN points, so nothing is sharded afterwards);
        # None -> IDRLNET_B200_DEVICE_SAMPLING; off by default because the point stream then differs from NumPy's
        if device_sampling is None:
            device_sampling = os.environ.get("IDRLNET_B200_DEVICE_SAMPLING", "0") not in ("0", "false", "False", "")
        self.device_sampling = bool(device_sampling)
        if self.device_sampling:
            from . import DEVICE
            if DEVICE.type != "cuda":
                raise RuntimeError("Solver(device_sampling=True) needs a CUDA device: the samplers are CUDA kernels")
            from .geo_utils import geo as _geo
            _geo.set_device_sampling(True, seed=sampling_seed)
            self.shard_points = False
        # measurement hook: {domain name: list} -> a (start, stop) CUDA-event pair around that domain's fused kernel is
        # appended at every gradient step (bench.py's roofline.kernel_ms)
        self.kernel_events: Dict[str, list] = {}
        self._move_nets_to_device()
        self.init_network_dirs = init_network_dirs if init_network_dirs else []
        self.init_load()
        self.pdes: List = [] if pdes is None else pdes
        pathlib.Path(self.network_dir).mkdir(parents=True, exist_ok=True)
        self.global_step = 0
        self.max_iter, self.save_freq, self.print_freq = max_iter, save_freq, print_freq
        try:
            self.parse_configure(**({"opt_config": opt_config} if opt_config is not None else {}),
                                 **({"schedule_config": schedule_config} if schedule_config is not None else {}))
        except Exception:
            logger.error("Optimizer configuration failed")
            raise
        if loading:
            try:
                self.load()
            except Exception:
                pass  # idrlnet/solver.py:116-120
        self._receivers: List[Receiver] = []
        self.sample_domains = sample_domains
        self.summary_dir = self.network_dir if summary_dir is None else summary_dir
        self.receivers = [callbacks.SummaryReceiver(self.summary_dir), callbacks.HandleResultReceiver(result_dir)]

    # ------------------------------------------------------------------ setup
    def _move_nets_to_device(self):
        from . import DEVICE
        for nn in self.netnodes:
            nn.net.to(DEVICE)

    @property
    def network_dir(self):
        return self._network_dir

    @network_dir.setter
    def network_dir(self, network_dir):
        self._network_dir = network_dir

    @property
    def sample_domains(self):
        return self._sample_domains

    @sample_domains.setter
    def sample_domains(self, sample_domains):
        self._sample_domains = sample_domains
        self.domain_losses = {d.name: d.loss_fn for d in sample_domains}
        self._generate_dict_index()
        self.generate_computation_pipeline()

    @property
    def trainable_parameters(self):
        groups = [{"params": list(nn.net.parameters())} for nn in self.netnodes
                  if not nn.is_reference and not nn.fixed]
        if not groups:
            logger.warning("No trainable parameters found!")
            return [torch.nn.parameter.Parameter(torch.tensor([0.0]), requires_grad=True)]
        return groups

    @property
    def summary_receiver(self):
        return self.receivers[0]

    def __str__(self):
        return ("nets: \n" + "".join(str(n) for n in self.netnodes) + "domains: \n"
                + "".join(str(d) for d in self.sample_domains) + "\noptimizer config:\n" + Optimizable.__str__(self))

    def set_param_ranges(self, param_ranges: Dict):
        for domain in self.sample_domains:
            domain.sample_fn.param_ranges = param_ranges

    def set_domain_parameter(self, domain_name: str, parameter_dict: dict):
        domain = self.get_sample_domain(domain_name)
        for key, value in parameter_dict.items():
            domain.sample_fn.__dict__[key] = value

    def get_domain_parameter(self, domain_name: str, parameter: str):
        return self.get_sample_domain(domain_name).sample_fn.__dict__[parameter]

    def get_sample_domain(self, name: str) -> DataNode:
        for d in self.sample_domains:
            if d.name == name:
                return d
        raise KeyError(f"domain {name} not exist!")

    def append_sample_domain(self, datanode):
        self.sample_domains = tuple(self.sample_domains) + (datanode,)

    def _generate_dict_index(self) -> None:
        self.invar_dict_index = {d.name: d.inputs for d in self.sample_domains}
        self.outvar_dict_index = {d.name: d.outputs for d in self.sample_domains}
        self.lambda_dict_index = {d.name: d.lambda_outputs for d in self.sample_domains}

    # ------------------------------------------------------- graph + fusion
    def generate_computation_pipeline(self):
        """One graph per domain (idrlnet/solver.py:214-228) and, where possible,
        one fused-kernel binding per domain."""
        samples = self.sample_variables_from_domains()
        in_var, _, _ = self.generate_in_out_dict(samples)
        self.vertex_pipelines = {}
        for name, var in in_var.items():
            logger.debug(f"Constructing computation graph for domain <{name}>")
            self.vertex_pipelines[name] = VertexTaskPipeline(self.netnodes + self.pdes, var, self.outvar_dict_index[name])
        self._plan_fusion(in_var)

    def _plan_fusion(self, in_var: DomainVariables) -> None:
        from . import DEVICE, fused
        self._fused_domains: Dict[str, "fused.FusedDomain"] = {}
        self._fused_owner: Dict[str, NetNode] = {}
        self._engines: Dict[int, "fused.StepEngine"] = {}
        self._packs = getattr(self, "_packs", {})
        want = self.fused_enabled
        if want is False or DEVICE.type != "cuda":
            if want:
                raise RuntimeError("Solver(fused=True) needs a CUDA device: the fused step has no CPU implementation")
            return
        exprs = {}
        for pde in self.pdes:
            if isinstance(pde, PdeNode):
                for sub in pde.sub_nodes:
                    if hasattr(sub.evaluate, "tf_eq"):
                        exprs[sub.name] = sub.evaluate.tf_eq
        specs = [cp.NetSpec(nn.name, tuple(nn.inputs), tuple(nn.outputs)) for nn in self.netnodes]
        by_name = {nn.name: nn for nn in self.netnodes}
        for dom in self.sample_domains:
            if not dom.outputs:
                continue
            try:
                cd = cp.compile_domain(list(dom.outputs), exprs, specs, list(in_var[dom.name].keys()))
                nn = by_name[cd.net.name]
                if nn.fixed or nn.require_no_grad:
                    raise cp.NotFusable("fixed / no-grad net")
                if isinstance(dom.sigma, torch.Tensor) and dom.sigma.requires_grad:
                    raise cp.NotFusable("trainable domain weight (var_sigma): the sigma leaf gets its grad from autograd")
                key = id(nn.net)
                if key not in self._packs:
                    self._packs[key] = fused.NetPack(nn.net, precision=self.precision)
                pack = self._packs[key]
                if pack.n_in != len(nn.inputs) or pack.n_out < len(nn.outputs):
                    raise cp.NotFusable("net shape does not match the node's inputs/outputs")
                fd = fused.FusedDomain(dom.name, cd, dom.loss_fn if isinstance(dom.loss_fn, str) else dom.loss_fn.name,
                                       float(dom.sigma))
                import ctypes as C
                from . import native as nv
                pack.refresh_effective()
                if not nv.lib().pinn_supported(C.byref(pack.mlp_struct()), C.byref(fd.jets)):
                    raise cp.NotFusable("no kernel for this net width / jet plan")
                self._fused_domains[dom.name] = fd
                self._fused_owner[dom.name] = nn
            except cp.NotFusable as e:
                logger.info(f"domain <{dom.name}>: generic path ({e})")
        if want and len(self._fused_domains) == 0:
            raise RuntimeError("Solver(fused=True): no domain could be fused")
        if self._fused_domains:
            logger.info("fused CUDA step for domains: " + ", ".join(self._fused_domains))

    # --------------------------------------------------------------- forward
    def forward_through_all_graph(self, invar_dict: DomainVariables,
                                  req_outvar_dict_index: Dict[str, List[str]]) -> DomainVariables:
        return {key: self.vertex_pipelines[key].forward_pipeline(invar_dict[key], req)
                for key, req in req_outvar_dict_index.items()}

    def generate_in_out_dict(self, samples: DomainVariables):
        def pick(index):
            return {d: Variables({k: v for k, v in var.items() if k in index[d]}) for d, var in samples.items()}
        return pick(self.invar_dict_index), pick(self.outvar_dict_index), pick(self.lambda_dict_index)

    def _sigma(self, domain_name: str, device):
        """The domain weight sigma: a python float when it is a constant (no per-step host-to-device copy),
        the tensor itself when it is trainable (`var_sigma`)."""
        sig = self.get_sample_domain(domain_name).sigma
        if isinstance(sig, torch.Tensor):
            return sig.to(device) if sig.requires_grad else float(sig)
        return float(sig)

    def sample_variables_from_domains(self) -> DomainVariables:
        return {d.name: d.sample() for d in self.sample_domains}

    # ------------------------------------------------------------------ train
    def solve(self):
        self.notify(self, message={Signal.SOLVE_START: "default"})
        while self.global_step < self.max_iter:
            loss = self.train_pipe()
            if self.global_step % self.print_freq == 0:
                logger.info("Iteration: {}, Loss: {}".format(self.global_step, float(loss)))
            if self.global_step % self.save_freq == 0:
                self.save()
        logger.info("Training Stage Ends")
        self.notify(self, message={Signal.SOLVE_END: "default"})

    def train_pipe(self):
        """Sample once, loss + gradients once, optimiser step once."""
        self.notify(self, message={Signal.TRAIN_PIPE_START: "defaults"})
        loss = None
        for opt in self.optimizers:
            if self.optimizer_config["optimizer"] == "LBFGS":
                def closure():
                    opt.zero_grad()
                    return self._loss_and_grads()
                loss = opt.step(closure)
            else:
                opt.zero_grad()
                loss = self._loss_and_grads()
                with _Range("optimizer"):
                    opt.step()
        if self.post_step_loss:
            with _Range("post_step_loss"):
                loss = self._loss_and_grads(grads=False)  # idrlnet/solver.py:341-344
        self.global_step += 1
        for scheduler in self.schedulers:
            scheduler.step(self.global_step)
        self.notify(self, message={Signal.TRAIN_PIPE_END: "defaults"})
        return loss

    def _shard(self, samples: DomainVariables) -> DomainVariables:
        d = _dist()
        if d is None or not self.shard_points:
            return samples
        r, w = d.get_rank(), d.get_world_size()
        return {name: Variables({k: (v[r::w] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] > 1 else v)
                                 for k, v in var.items()}) for name, var in samples.items()}

    def _loss_and_grads(self, grads: bool = True) -> torch.Tensor:
        with _Range("sample"):
            samples = self._shard(self.sample_variables_from_domains())
            in_var, true_out, lambda_out = self.generate_in_out_dict(samples)
        fused_names = [n for n in self._fused_domains if len(true_out[n]) > 0]
        generic = {n: idx for n, idx in self.outvar_dict_index.items() if n not in fused_names and len(true_out[n]) > 0}
        comps: Dict[str, torch.Tensor] = {}
        gen_loss = None
        # -- generic domains: graph forward + torch loss ------------------------
        if generic:
            with _Range("generic_graph"):
                pred = self.forward_through_all_graph(in_var, generic)
                comps.update(self._loss_components(in_var, pred, {n: true_out[n] for n in generic},
                                                   {n: lambda_out[n] for n in generic}))
        # -- fused domains: one kernel per domain --------------------------------
        engines = self._bind_fused(fused_names, in_var, true_out, lambda_out)
        dist = _dist()
        for eng in engines:
            timed = None
            if grads and self.kernel_events:
                for i, fd in enumerate(eng.domains):
                    if fd.name in self.kernel_events:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        self.kernel_events[fd.name].append(ev)
                        timed = (i, ev[0], ev[1])
                        break
            with _Range("fused_step"):
                eng.launch(loss_only=not grads, timed=timed)  # the post-update re-evaluation needs no reverse pass
            if dist is not None and not eng.reduced:  # (narrow nets: finalize already summed over the ranks)
                with _Range("all_reduce"):
                    dist.all_reduce(eng.flat if grads else eng.losses)
            for fd, l in zip(eng.domains, eng.losses):
                comps[f"{fd.name}_loss"] = l
        ordered = {f"{d.name}_loss": comps[f"{d.name}_loss"] for d in self.sample_domains if f"{d.name}_loss" in comps}
        self.loss_component = Variables(ordered)
        self.notify(self, message={Signal.BEFORE_COMPUTE_LOSS: {**self.loss_component}})
        if len(engines) == 1 and not generic and list(ordered) == [f"{fd.name}_loss" for fd in engines[0].domains]:
            loss = engines[0].total_loss()  # sum_d sigma_d * loss_d as ONE dot product (was 2 tiny kernels per domain)
        else:
            loss = sum(self._sigma(k[:-5], v.device) * v for k, v in ordered.items())
        self.notify(self, message={Signal.AFTER_COMPUTE_LOSS: {**self.loss_component, **{"total_loss": loss}}})
        if not grads:
            return loss.detach() if isinstance(loss, torch.Tensor) else loss
        self.notify(self, message={Signal.BEFORE_BACKWARD: "defaults"})
        if generic:
            gen_loss = sum(self._sigma(n, comps[f"{n}_loss"].device) * comps[f"{n}_loss"]
                           for n in generic)
            if isinstance(gen_loss, torch.Tensor) and gen_loss.requires_grad:
                gen_loss.backward()
                if dist is not None:
                    self._all_reduce_param_grads(dist)
        for eng in engines:
            eng.pack.scatter_grad(eng.grad, accumulate=bool(generic) or len(engines) > 1)
        return loss.detach() if isinstance(loss, torch.Tensor) else loss

    def _all_reduce_param_grads(self, dist) -> None:
        """SUM of the generic-path `.grad`s over the ranks: every Parameter once (a net shared through
        `get_shared_net_node` shows up in two NetNodes with the same Parameter objects), flattened into one
        buffer -> one collective, as on the fused path."""
        seen, grads = set(), []
        for nn in self.netnodes:
            if nn.is_reference:
                continue
            for p in nn.net.parameters():
                if p.grad is not None and id(p) not in seen:
                    seen.add(id(p))
                    grads.append(p.grad)
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _bind_fused(self, names, in_var, true_out, lambda_out):
        from . import fused
        groups: Dict[int, List] = {}
        for n in names:
            fd = self._fused_domains[n]
            pts = in_var[n]
            lambdas = {k[len("lambda_"):]: v for k, v in lambda_out[n].items()}
            for k in list(true_out[n].keys()):
                if k not in lambdas and ("lambda_" + k) in pts:  # idrlnet/solver.py:371-378
                    lambdas[k] = pts["lambda_" + k]
            fd.sigma = float(self.get_sample_domain(n).sigma)  # read every step, as idrlnet/solver.py:394 does
            fd.bind(pts, {k: v for k, v in true_out[n].to_torch_tensor_().items()}, lambdas, pts.get("area"))
            groups.setdefault(id(self._fused_owner[n].net), []).append(fd)
        engines = []
        for key, fds in groups.items():
            sig = tuple(fd.name for fd in fds)
            eng = self._engines.get(key)
            if eng is None or getattr(eng, "_sig", None) != sig:
                eng = fused.StepEngine(self._packs[key], fds)
                eng._sig = sig
                self._engines[key] = eng
                if _dist() is not None and os.environ.get("IDRLNET_B200_PEER_ALLREDUCE", "1") != "0":
                    eng.enable_peer_allreduce()  # collective; every rank builds the same engines in the same order
            else:
                eng.set_domains(fds)
            engines.append(eng)
        return engines

    def _loss_components(self, in_var, pred_out_sample, true_out, lambda_out) -> Dict[str, torch.Tensor]:
        """Per-domain weighted losses of the generic path (idrlnet/solver.py:362-390)."""
        comps = {}
        for name, tgt in true_out.items():
            if len(tgt) == 0:
                continue
            diff = pred_out_sample[name] - tgt.to_torch_tensor_()
            diff.update(lambda_out[name])
            diff.update(area=in_var[name]["area"])
            for k in list(diff.keys()):
                if "lambda_" + k in in_var[name]:
                    diff["lambda_" + k] = in_var[name]["lambda_" + k]
            comps.update(diff.weighted_loss(f"{name}_loss", loss_function=self.domain_losses[name]))
        return comps

    def compute_loss(self, in_var: DomainVariables, pred_out_sample: DomainVariables, true_out: DomainVariables,
                     lambda_out: DomainVariables) -> torch.Tensor:
        """Total loss of already-forwarded domains (idrlnet/solver.py:353-408)."""
        comps = self._loss_components(in_var, pred_out_sample, true_out, lambda_out)
        self.loss_component = Variables(comps)
        self.notify(self, message={Signal.BEFORE_COMPUTE_LOSS: {**self.loss_component}})
        loss = sum(self._sigma(k[:-5], v.device) * v for k, v in comps.items())
        self.notify(self, message={Signal.AFTER_COMPUTE_LOSS: {**self.loss_component, **{"total_loss": loss}}})
        return loss

    def infer_step(self, domain_attr: Dict[str, List[str]]) -> DomainVariables:
        samples = self.sample_variables_from_domains()
        in_var, _, _ = self.generate_in_out_dict(samples)
        return self.forward_through_all_graph(in_var, domain_attr)

    def residual_top_k(self, domain_name: str, residual: str, k: int) -> Variables:
        """Adaptive re-sampling on the device: sample `domain_name`, evaluate the
        PdeNode output `residual` at every point with the forward-only kernels and return the k
        points with the largest |residual| -- every sampled column of those points, the residual
        itself and their `index` in the sample, ordered by descending |residual|.

        The reference's recipe (examples/allen_cahn/allen_cahn.py:128-149) is `infer_step` +
        `np.argsort` on the host; the selection here is identical (ties by ascending index).
        With torch.distributed initialised every rank scores its shard and the per-rank winners
        are all-gathered and re-selected, so all ranks return the same k points."""
        from . import DEVICE, fused
        if DEVICE.type != "cuda":
            raise RuntimeError("residual_top_k runs on the CUDA kernels; there is no CPU implementation")
        dom = self.get_sample_domain(domain_name)
        sample = self._shard({domain_name: dom.sample()})[domain_name]
        points = Variables({key: v for key, v in sample.items() if key in self.invar_dict_index[domain_name]})
        points.to_torch_tensor_()
        exprs = {sub.name: sub.evaluate.tf_eq for pde in self.pdes if isinstance(pde, PdeNode) for sub in pde.sub_nodes
                 if hasattr(sub.evaluate, "tf_eq")}
        specs = [cp.NetSpec(nn.name, tuple(nn.inputs), tuple(nn.outputs)) for nn in self.netnodes]
        cd = cp.compile_domain([residual], exprs, specs, list(points.keys()))
        nn = next(n for n in self.netnodes if n.name == cd.net.name)
        key = id(nn.net)
        if key not in self._packs:
            self._packs[key] = fused.NetPack(nn.net, precision=self.precision)
        fd = fused.FusedDomain(domain_name, cd)
        fd.bind(points, {residual: 0.0}, {}, 1.0)
        r = fused.domain_residuals(self._packs[key], fd)[0]
        n = r.numel()
        idx, val = fused.topk_abs(r, min(k, n))
        picked = {name: v.detach().reshape(n, -1)[idx] for name, v in points.items()
                  if isinstance(v, torch.Tensor) and v.numel() >= n and v.shape[0] == n}
        picked[residual] = val.reshape(-1, 1)
        picked["index"] = idx.reshape(-1, 1)
        d = _dist()
        if d is not None and d.get_world_size() > 1:
            world = d.get_world_size()
            picked["index"] = picked["index"] * world + d.get_rank()  # position in the unsharded sample
            gathered = {}
            for name, v in picked.items():
                parts = [torch.empty_like(v) for _ in range(world)]
                d.all_gather(parts, v.contiguous())
                gathered[name] = torch.cat(parts, dim=0)
            sel, _ = fused.topk_abs(gathered[residual], min(k, gathered[residual].numel()))
            picked = {name: v[sel] for name, v in gathered.items()}
        return Variables(picked)

    # ------------------------------------------------------------- checkpoint
    def save(self):
        """`model.ckpt`: {<net>_dict, optimizer_<i>_dict, global_step} (idrlnet/solver.py:425-436)."""
        d = _dist()
        if d is not None and d.get_rank() != 0:
            return
        path = os.path.join(self.network_dir, "model.ckpt")
        logger.info("save to path: {}".format(os.path.abspath(path)))
        save_dict = {f"{nn.name}_dict": nn.state_dict() for nn in self.netnodes if not nn.is_reference}
        for i, opt in enumerate(self.optimizers):
            save_dict[f"optimizer_{i}_dict"] = opt.state_dict()
        save_dict["global_step"] = self.global_step
        torch.save(save_dict, path)

    def init_load(self):
        from . import DEVICE
        for network_dir in self.init_network_dirs:
            save_dict = torch.load(os.path.join(network_dir, "model.ckpt"), map_location=DEVICE)
            for nn in self.netnodes:
                if f"{nn.name}_dict" in save_dict and not nn.is_reference:
                    nn.load_state_dict(save_dict[f"{nn.name}_dict"])
                    logger.info(f"Successfully loading initialization {nn.name}.")

    def load(self):
        from . import DEVICE
        save_dict = torch.load(os.path.join(self.network_dir, "model.ckpt"), map_location=DEVICE)
        for i, opt in enumerate(self.optimizers):
            opt.load_state_dict(save_dict[f"optimizer_{i}_dict"])
        self.global_step = save_dict["global_step"]
        for nn in self.netnodes:
            if f"{nn.name}_dict" in save_dict and not nn.is_reference:
                nn.load_state_dict(save_dict[f"{nn.name}_dict"])
                logger.info(f"Successfully loading {nn.name}.")

    def configure_optimizers(self):
        opt = self.optimizer_config["optimizer"]
        if isinstance(opt, str) and opt in Optimizable.OPTIMIZER_MAP:
            opt = Optimizable.OPTIMIZER_MAP[opt](self.trainable_parameters,
                                                 **{k: v for k, v in self.optimizer_config.items() if k != "optimizer"})
        elif not isinstance(opt, Callable):
            raise NotImplementedError("The optimizer is not implemented. You may use one of the following optimizer:\n"
                                      + "\n".join(Optimizable.OPTIMIZER_MAP.keys())
                                      + '\n Example: opt_config=dict(optimizer="Adam", lr=1e-3)')
        sch = self.schedule_config["scheduler"]
        if isinstance(sch, str) and sch in Optimizable.SCHEDULE_MAP:
            sch = Optimizable.SCHEDULE_MAP[sch](opt, **{k: v for k, v in self.schedule_config.items() if k != "scheduler"})
        elif not isinstance(sch, Callable):
            raise NotImplementedError("The scheduler is not implemented. You may use one of the following scheduler:\n"
                                      + "\n".join(Optimizable.SCHEDULE_MAP.keys())
                                      + '\n Example: schedule_config=dict(scheduler="ExponentialLR", gamma=0.999')
        self.optimizers = [opt]
        self.schedulers = [sch]
