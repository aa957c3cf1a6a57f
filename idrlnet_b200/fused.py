"""Torch-facing host layer over the C ABI (include/pinn_b200.h).

PyTorch is used here for device memory, streams and (multi-GPU)
`torch.distributed`; all per-point arithmetic happens in libpinn_b200.so.

* `NetPack`      -- the parameters of one MLP (`idrlnet/architecture/mlp.py:18-76`)
                    as the kernels see them: effective W/b per layer (weight-norm
                    resolved by a kernel), one flat gradient buffer.
* `JetFunction`  -- autograd op "MLP + input-derivative jets" for the generic
                    path (replaces `idrlnet/net.py:14-36` +
                    `idrlnet/variable.py:161-208` when a residual cannot be fused).
* `FusedDomain`  -- one sampling domain compiled for `pinn_step_fused`.
* `StepEngine`   -- one training step over several domains sharing a net:
                    forward jets + residual + loss + reverse pass + gradient
                    reduction (replaces `idrlnet/solver.py:329-338`).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import torch

from . import compiler as cp
from . import native as nv


def default_precision() -> str:
    """Arithmetic of the wide-net (H > 32) layer GEMMs.  Default "fp32x3": tcgen05 tensor cores with 3xTF32 operand
    splitting -- FP32-grade products (whole-gradient norm-wise error <= 1e-5 against the FP64 oracle, the same bar the
    FP32 FMA kernels are held to in tests/test_gpu_wide.py) at 1.4-3.8x the speed of "fp32" (FP32 FMA everywhere).
    "tf32": tensor cores with 10-bit multiplicands, stated tolerance 1e-2 (on-chip fused kernel where it applies, the
    mode the >= 20x figures on ldc are quoted in); "tf32_streamed": the layer-streamed TF32 kernels.  Narrow nets
    (H <= 32) always run FP32 FMA.  Set IDRLNET_B200_PRECISION or pass precision= explicitly."""
    import os
    return os.environ.get("IDRLNET_B200_PRECISION", "fp32x3").lower()


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"idrlnet_b200: {what} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"idrlnet_b200: {what} must be float32, got {t.dtype}")


def _linear_layers(module: torch.nn.Module) -> List[torch.nn.Module]:
    lins = [m for m in module.modules() if isinstance(m, torch.nn.Linear)]
    if not lins:
        raise cp.NotFusable("net has no Linear layers")
    return lins


def _activation_name(module: torch.nn.Module) -> str:
    """Name of the single activation used between the Linear layers."""
    names = set()
    for m in module.modules():
        n = type(m).__name__.lower()
        if n in ("tanh",):
            names.add("tanh")
        elif n in ("silu", "swish"):
            names.add("silu")
        elif n in ("sin",):
            names.add("sin")
        elif n in ("relu", "selu", "sigmoid", "poly", "leaky_relu", "leakyrelu", "gelu", "elu", "softplus"):
            names.add(n)
    if len(names) != 1:
        raise cp.NotFusable(f"net must use exactly one activation kind, found {sorted(names)}")
    act = names.pop()
    if act not in nv.ACT:
        raise cp.NotFusable(f"activation {act} has no fused kernel")
    return act


class NetPack:
    """Kernel view of an MLP: Linear stack, uniform hidden width, one activation."""

    def __init__(self, module: torch.nn.Module, act: Optional[str] = None, precision: Optional[str] = None):
        self.module = module
        self.precision = precision or default_precision()
        if self.precision not in nv.PRECISION:
            raise ValueError(f"precision must be one of {sorted(nv.PRECISION)}")
        self.layers = _linear_layers(module)
        self.act = act or _activation_name(module)
        if len(self.layers) < 2:
            raise cp.NotFusable("need at least one hidden layer")
        self.n_in = self.layers[0].in_features
        self.n_out = self.layers[-1].out_features
        self.width = self.layers[0].out_features
        self.n_hidden = len(self.layers) - 1
        for i, lay in enumerate(self.layers):
            want_in = self.n_in if i == 0 else self.width
            want_out = self.n_out if i == self.n_hidden else self.width
            if lay.in_features != want_in or lay.out_features != want_out:
                raise cp.NotFusable("hidden layers must all have the same width")
            if lay.bias is None:
                raise cp.NotFusable("Linear layers without bias are not supported")
        if self.n_in > nv.MAX_IN or self.n_out > nv.MAX_OUT or len(self.layers) > nv.MAX_LAYERS:
            raise cp.NotFusable("net too large for the descriptors")
        self.weight_norm = [hasattr(l, "weight_g") and hasattr(l, "weight_v") for l in self.layers]
        # Siren (idrlnet/architecture/mlp.py:79-137): hidden layer i computes sin(omega_i * (W x + b)); folded into the
        # effective parameters the kernels read, W_eff = omega_i * W, b_eff = omega_i * b (the `sin` jets exist already)
        self.scale = [1.0] * len(self.layers)
        if hasattr(module, "first_omega") and hasattr(module, "omega"):
            for i in range(self.n_hidden):
                self.scale[i] = float(module.first_omega if i == 0 else module.omega)
            if any(self.weight_norm):
                raise cp.NotFusable("omega-scaled layers with weight-norm are not supported")
        if any(p.dtype != torch.float32 for l in self.layers for p in l.parameters()):
            raise cp.NotFusable("the kernels compute in FP32; a net of another dtype (e.g. an FP64 tie-breaker run) "
                                "takes the torch path")
        dev = self.layers[0].bias.device
        if dev.type != "cuda":
            raise RuntimeError("idrlnet_b200: the net must live on a CUDA device (there is no CPU path)")
        self.device = dev
        self.sizes = []
        for l in self.layers:
            self.sizes += [l.out_features * l.in_features, l.out_features]
        self.n_params = sum(self.sizes)
        # effective-weight buffers for weight-norm layers (plain layers are read in place)
        self._w_eff = [torch.empty(l.out_features, l.in_features, device=dev, dtype=torch.float32) if (wn or sc != 1.0) else None
                       for l, wn, sc in zip(self.layers, self.weight_norm, self.scale)]
        self._b_eff = [torch.empty(l.out_features, device=dev, dtype=torch.float32) if sc != 1.0 else None
                       for l, sc in zip(self.layers, self.scale)]
        self._grad_wn: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._wn_fwd = None   # (pointer key, pinn_wn_layer array) of the batched weight-norm launches
        self._wn_bwd = None
        self._mlp = None
        self._mlp_key = None
        self._probe()

    def _probe(self) -> None:
        """The kernels evaluate `act(W h + b)` layer by layer.  A module with the right layers
        but another forward (e.g. Siren's omega scaling, skip connections) must not be fused:
        compare the module with that formula on a few random points."""
        fn = {"tanh": torch.tanh, "silu": torch.nn.functional.silu, "sin": torch.sin}[self.act]
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(1234)
            x = (torch.rand(16, self.n_in, generator=g, device="cpu") * 2 - 1).to(self.device)  # explicit device: the
            # reference's use_gpu() makes CUDA the default tensor type (idrlnet/__init__.py:29)
            want = self.module(x)
            h = x
            for i, l in enumerate(self.layers):
                W = torch._weight_norm(l.weight_v, l.weight_g, 0) if self.weight_norm[i] else l.weight
                h = torch.nn.functional.linear(h, W, l.bias) * self.scale[i]
                if i < self.n_hidden:
                    h = fn(h)
        if want.shape != h.shape or not torch.allclose(want, h, rtol=1e-4, atol=1e-5 * float(h.abs().max() + 1e-30)):
            raise cp.NotFusable("module.forward is not a plain Linear/activation stack")

    # -- parameters ---------------------------------------------------------
    def raw_params(self, i: int):
        l = self.layers[i]
        if self.weight_norm[i]:
            return l.weight_g, l.weight_v, l.bias
        return l.weight, l.bias

    def effective(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        l = self.layers[i]
        if self.scale[i] != 1.0:
            return self._w_eff[i], self._b_eff[i]
        return (self._w_eff[i] if self.weight_norm[i] else l.weight.data), l.bias.data

    def differentiable_effective(self) -> List[torch.Tensor]:
        """[W_0, b_0, W_1, b_1, ...] as autograd functions of the module's own parameters (generic path)."""
        wb = []
        for i, l in enumerate(self.layers):
            W = torch._weight_norm(l.weight_v, l.weight_g, 0) if self.weight_norm[i] else l.weight
            b = l.bias
            if self.scale[i] != 1.0:
                W, b = W * self.scale[i], b * self.scale[i]
            wb += [W, b]
        return wb

    def refresh_effective(self) -> None:
        """W = g * v / ||v|| for the weight-norm layers (one launch for all of them); omega * (W, b) for Siren layers."""
        scaled = [i for i, sc in enumerate(self.scale) if sc != 1.0]
        if scaled:
            with torch.no_grad():
                torch._foreach_mul_  # (availability check: torch >= 1.9)
                outs = [self._w_eff[i] for i in scaled] + [self._b_eff[i] for i in scaled]
                srcs = [self.layers[i].weight.data for i in scaled] + [self.layers[i].bias.data for i in scaled]
                torch._foreach_copy_(outs, srcs)
                torch._foreach_mul_(outs, [self.scale[i] for i in scaled] * 2)
        idx = [i for i, wn in enumerate(self.weight_norm) if wn]
        if not idx:
            return
        ptrs = tuple((self.layers[i].weight_g.data.data_ptr(), self.layers[i].weight_v.data.data_ptr()) for i in idx)
        if self._wn_fwd is None or self._wn_fwd[0] != ptrs:
            arr = (nv.PinnWnLayer * len(idx))()
            for a, i in zip(arr, idx):
                l = self.layers[i]
                a.g, a.v, a.W = l.weight_g.data.data_ptr(), l.weight_v.data.data_ptr(), self._w_eff[i].data_ptr()
                a.out_dim, a.in_dim = l.out_features, l.in_features
            self._wn_fwd = (ptrs, arr)
        nv.check(nv.lib().pinn_weight_norm_forward_batch(self._wn_fwd[1], len(idx), _stream_ptr()))

    def mlp_struct(self) -> nv.PinnMlp:
        ptrs = tuple(t.data_ptr() for i in range(len(self.layers)) for t in self.effective(i))
        if self._mlp is None or self._mlp_key != ptrs:
            for i in range(len(self.layers)):
                for t in self.effective(i):
                    _require_cuda(t, "net parameter")
                    if not t.is_contiguous():
                        raise RuntimeError("idrlnet_b200: net parameters must be contiguous")
            self._mlp = nv.make_mlp(self.n_in, self.n_out, self.n_hidden, self.width, self.act, ptrs[0::2], ptrs[1::2],
                                    self.precision)
            self._mlp_key = ptrs
        return self._mlp

    # -- gradients ----------------------------------------------------------
    def scatter_grad(self, flat: torch.Tensor, accumulate: bool = False) -> None:
        """flat = d loss / d(effective W, b) in pinn_param_count order -> `.grad`
        of the module's own parameters (through weight-norm where present, one launch for all layers)."""
        idx = [i for i, wn in enumerate(self.weight_norm) if wn]
        if idx:
            for i in idx:
                if i not in self._grad_wn:
                    l = self.layers[i]
                    self._grad_wn[i] = (torch.empty_like(l.weight_g.data), torch.empty_like(l.weight_v.data))
            key = (flat.data_ptr(),) + tuple((self.layers[i].weight_g.data.data_ptr(),
                                              self.layers[i].weight_v.data.data_ptr()) for i in idx)
            if self._wn_bwd is None or self._wn_bwd[0] != key:
                arr = (nv.PinnWnLayer * len(idx))()
                offs, off = {}, 0
                for i in range(len(self.layers)):
                    offs[i] = off
                    off += self.sizes[2 * i] + self.sizes[2 * i + 1]
                for a, i in zip(arr, idx):
                    l = self.layers[i]
                    dg, dv = self._grad_wn[i]
                    a.g, a.v = l.weight_g.data.data_ptr(), l.weight_v.data.data_ptr()
                    a.dW = flat.data_ptr() + 4 * offs[i]
                    a.dg, a.dv = dg.data_ptr(), dv.data_ptr()
                    a.out_dim, a.in_dim = l.out_features, l.in_features
                self._wn_bwd = (key, arr)
            nv.check(nv.lib().pinn_weight_norm_backward_batch(self._wn_bwd[1], len(idx), _stream_ptr()))
        off = 0
        for i, l in enumerate(self.layers):
            nW, nb = self.sizes[2 * i], self.sizes[2 * i + 1]
            dW = flat[off:off + nW].view(l.out_features, l.in_features)
            db = flat[off + nW:off + nW + nb]
            off += nW + nb
            if self.scale[i] != 1.0:  # d/dW = omega * d/dW_eff (in place: the flat buffer is rewritten by the next step)
                dW.mul_(self.scale[i])
                db.mul_(self.scale[i])
            if self.weight_norm[i]:
                dg, dv = self._grad_wn[i]
                _set_grad(l.weight_g, dg, accumulate)
                _set_grad(l.weight_v, dv, accumulate)
            else:
                _set_grad(l.weight, dW, accumulate)
            _set_grad(l.bias, db, accumulate)


def _set_grad(p: torch.nn.Parameter, g: torch.Tensor, accumulate: bool) -> None:
    if accumulate and p.grad is not None:
        p.grad.add_(g)
    else:
        p.grad = g.clone() if accumulate else g


# --------------------------------------------------------------------------- #
class JetFunction(torch.autograd.Function):
    """jets[c*n_out + o, n] of an MLP given as effective (W, b) tensors.

    Differentiable w.r.t. the weights (not w.r.t. the coordinates: the input
    derivatives ARE the jet channels)."""

    @staticmethod
    def forward(ctx, plan: cp.JetPlan, act: str, precision: str, coords: Tuple[torch.Tensor, ...], *wb: torch.Tensor):
        lib = nv.lib()
        n_layers = len(wb) // 2
        Ws, bs = wb[0::2], wb[1::2]
        for t in list(coords) + list(wb):
            _require_cuda(t, "JetFunction argument")
        n = coords[0].numel()
        cols = [c.detach().reshape(-1).contiguous() for c in coords]
        wbc = [t.detach().contiguous() for t in wb]
        mlp = nv.make_mlp(len(cols), Ws[-1].shape[0], n_layers - 1, Ws[0].shape[0], act,
                          [t.data_ptr() for t in wbc[0::2]], [t.data_ptr() for t in wbc[1::2]], precision)
        jets = nv.make_jets(plan.first_in, plan.n_second, plan.n_mixed, plan.n_high)
        out = torch.empty(plan.n_channels * Ws[-1].shape[0], n, device=cols[0].device, dtype=torch.float32)
        ws_bytes = lib.pinn_workspace_bytes(C.byref(mlp), C.byref(jets), 1)
        ws = torch.empty(max(ws_bytes // 4, 1), device=cols[0].device, dtype=torch.float32)
        nv.check(lib.pinn_jet_forward(C.byref(mlp), C.byref(jets), n, nv.ptr_array([c.data_ptr() for c in cols]),
                                      out.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr()))
        ctx.plan, ctx.act, ctx.cols, ctx.wbc, ctx.precision = plan, act, cols, wbc, precision
        return out

    @staticmethod
    def backward(ctx, out_bar: torch.Tensor):
        lib = nv.lib()
        wbc, cols, plan = ctx.wbc, ctx.cols, ctx.plan
        Ws = wbc[0::2]
        n = cols[0].numel()
        mlp = nv.make_mlp(len(cols), Ws[-1].shape[0], len(Ws) - 1, Ws[0].shape[0], ctx.act,
                          [t.data_ptr() for t in wbc[0::2]], [t.data_ptr() for t in wbc[1::2]], ctx.precision)
        jets = nv.make_jets(plan.first_in, plan.n_second, plan.n_mixed, plan.n_high)
        ws_bytes = lib.pinn_workspace_bytes(C.byref(mlp), C.byref(jets), 1)
        ws = torch.empty(ws_bytes // 4, device=cols[0].device, dtype=torch.float32)
        ob = out_bar.contiguous()
        st = _stream_ptr()
        nv.check(lib.pinn_jet_backward(C.byref(mlp), C.byref(jets), n, nv.ptr_array([c.data_ptr() for c in cols]),
                                       ob.data_ptr(), ws.data_ptr(), ws_bytes, 0, st))
        flat = torch.empty(lib.pinn_param_count(C.byref(mlp)), device=ob.device, dtype=torch.float32)
        nv.check(lib.pinn_step_finalize(C.byref(mlp), C.byref(jets), ws.data_ptr(), ws_bytes, 1, flat.data_ptr(), 0,
                                        None, st))
        grads, off = [], 0
        for t in wbc:
            grads.append(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        return (None, None, None, None, *grads)


def jet_mlp(plan: cp.JetPlan, act: str, coords: Sequence[torch.Tensor], weights_and_biases: Sequence[torch.Tensor],
            precision: Optional[str] = None):
    return JetFunction.apply(plan, act, precision or default_precision(), tuple(coords), *weights_and_biases)


# --------------------------------------------------------------------------- #
Value = Union[torch.Tensor, float, int, None]


class FusedDomain:
    """A sampling domain bound to device columns, ready for `pinn_step_fused`."""

    def __init__(self, name: str, compiled: cp.CompiledDomain, loss: str = "square", sigma: float = 1.0):
        self.name = name
        self.compiled = compiled
        self.loss = loss
        self.sigma = float(sigma)
        self.jets = nv.make_jets(compiled.plan.first_in, compiled.plan.n_second, compiled.plan.n_mixed, compiled.plan.n_high)
        self._struct: Optional[nv.PinnDomain] = None
        self._keep: List[torch.Tensor] = []
        self.n_points = 0

    def bind(self, columns: Mapping[str, torch.Tensor], targets: Mapping[str, Value],
             lambdas: Mapping[str, Value], area: Value) -> None:
        """columns[key] = sampled [N,1] / [N] tensors (coords and aux symbols)."""
        keep: List[torch.Tensor] = []

        def col(t: torch.Tensor) -> torch.Tensor:
            _require_cuda(t, f"domain {self.name} column")
            t = t.detach().reshape(-1)
            if not t.is_contiguous():
                t = t.contiguous()
            keep.append(t)
            return t

        cd = self.compiled
        need = list(cd.net.inputs) + list(cd.aux_keys)
        cols = {k: col(columns[k]) for k in need}
        n = cols[cd.net.inputs[0]].numel()
        for k, t in cols.items():
            if t.numel() != n:
                raise RuntimeError(f"domain {self.name}: column {k} has {t.numel()} points, expected {n}")

        def val(v: Value, what: str):
            if isinstance(v, torch.Tensor):
                if v.numel() == 1 and n != 1:
                    return float(v)
                t = col(v)
                if t.numel() != n:
                    raise RuntimeError(f"domain {self.name}: {what} has {t.numel()} points, expected {n}")
                return cp.Ptr(t.data_ptr())
            return v if v is None else float(v)

        tg = {e.name: val(targets[e.name], f"target {e.name}") for e in cd.equations}
        lm = {e.name: val(lambdas.get(e.name), f"lambda_{e.name}") for e in cd.equations}
        ar = val(area, "area")
        if ar is None:
            raise RuntimeError(f"domain {self.name}: no `area` (idrlnet/solver.py:369 requires it)")
        self._struct = cp.build_domain_struct(cd, n, lambda k: cols[k].data_ptr(), tg, lm, ar, self.loss, self.sigma)
        self._keep = keep
        self.n_points = n

    @property
    def struct(self) -> nv.PinnDomain:
        if self._struct is None:
            raise RuntimeError(f"domain {self.name} is not bound to data")
        return self._struct


def topk_abs(values: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(indices int64 [k], values [k]) of the k largest |values|, descending, ties by ascending
    index -- `np.argsort(-np.abs(v), kind="stable")[:k]` on the device (pinn_topk_abs)."""
    _require_cuda(values, "topk_abs input")
    v = values.detach().reshape(-1).to(torch.float32).contiguous()
    lib = nv.lib()
    ws_bytes = lib.pinn_topk_workspace_bytes(k)
    if ws_bytes < 0:
        raise nv.PinnError(int(ws_bytes), f"top-k supports 1 <= k <= 4096, got {k}")
    ws = torch.empty(ws_bytes, device=v.device, dtype=torch.uint8)
    idx = torch.empty(k, device=v.device, dtype=torch.int64)
    val = torch.empty(k, device=v.device, dtype=torch.float32)
    nv.check(lib.pinn_topk_abs(v.data_ptr(), v.numel(), k, idx.data_ptr(), val.data_ptr(), ws.data_ptr(), ws_bytes,
                               _stream_ptr()))
    return idx, val


def domain_residuals(pack: NetPack, domain: FusedDomain) -> torch.Tensor:
    """[n_eq, N] values of the domain's equations at its bound points, forward only
    (pinn_residual_forward): what `Solver.infer_step` returns for the PdeNode outputs."""
    lib = nv.lib()
    pack.refresh_effective()
    mlp = pack.mlp_struct()
    n = domain.n_points
    out = torch.empty((len(domain.compiled.equations), n), device=pack.device, dtype=torch.float32)
    if n == 0:
        return out
    ws_bytes = lib.pinn_residual_workspace_bytes(C.byref(mlp), C.byref(domain.jets), n)
    if ws_bytes < 0:
        raise nv.PinnError(int(ws_bytes), "no kernel for this net / jet plan")
    ws = torch.empty(ws_bytes // 4 + 1, device=pack.device, dtype=torch.float32)
    nv.check(lib.pinn_residual_forward(C.byref(mlp), C.byref(domain.jets), C.byref(domain.struct), out.data_ptr(),
                                       ws.data_ptr(), ws_bytes, _stream_ptr()))
    return out


class StepEngine:
    """One net, several fused domains: gradient of sum_d sigma_d * loss_d."""

    def __init__(self, pack: NetPack, domains: Sequence[FusedDomain]):
        self.pack = pack
        self.domains = list(domains)
        lib = nv.lib()
        pack.refresh_effective()
        mlp = pack.mlp_struct()
        self.ws_bytes = 0
        for d in self.domains:
            if not lib.pinn_supported(C.byref(mlp), C.byref(d.jets)):
                raise cp.NotFusable(f"no fused kernel for net width {pack.width} / jets of domain {d.name}")
            self.ws_bytes = max(self.ws_bytes, lib.pinn_workspace_bytes(C.byref(mlp), C.byref(d.jets), len(self.domains)))
        self.ws = torch.empty(max(self.ws_bytes // 4, 1), device=pack.device, dtype=torch.float32)
        # flat gradient with the per-domain losses appended: ONE buffer to all-reduce
        self.flat = torch.zeros(pack.n_params + len(self.domains), device=pack.device, dtype=torch.float32)
        # loss-only evaluations (the reference's post-update re-evaluation) write into their own buffer: `.grad`
        # of the module parameters may be views of `flat` and must survive until the next gradient step
        self._flat_eval: Optional[torch.Tensor] = None
        self._last = self.flat
        self._sigma_vals = [d.sigma for d in self.domains]
        self.sigmas = torch.tensor(self._sigma_vals, device=pack.device, dtype=torch.float32)
        # multi-GPU: finalize + all-reduce as ONE kernel over the ranks' symmetric memory (enable_peer_allreduce)
        self._peers: Optional[nv.PinnPeers] = None
        self._peer_keep = None
        self.reduced = False  # the most recent launch already summed `flat` over the ranks

    def enable_peer_allreduce(self) -> bool:
        """Under torch.distributed (NCCL, one node): let pinn_step_finalize_allreduce produce the gradient summed over the
        ranks -- the reduction kernel writes its values into every peer's symmetric region over NVLink, signals, waits
        and adds in rank order -- instead of pinn_step_finalize followed by an NCCL all-reduce of 6 KB.  COLLECTIVE: every
        rank must call it at the same point.  Returns False (and keeps the NCCL path) for wide nets, other backends or when
        symmetric memory is unavailable on any rank."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2 or dist.get_backend() != "nccl":
            return False
        world, rank = dist.get_world_size(), dist.get_rank()
        lib = nv.lib()
        self.pack.refresh_effective()
        mlp = self.pack.mlp_struct()
        nbytes = int(lib.pinn_allreduce_region_bytes(C.byref(mlp), C.byref(self.domains[0].jets), len(self.domains), world))
        ok, region, hdl = nbytes > 0, None, None
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm
                region = symm.empty(nbytes // 4, dtype=torch.float32, device=self.pack.device)
                region.zero_()
                hdl = symm.rendezvous(region, dist.group.WORLD)
                ok = len(hdl.buffer_ptrs) == world and hdl.rank == rank
            except Exception as e:  # no symmetric memory on this system: the NCCL path stays
                import logging
                logging.getLogger("idrlnet_b200").info("peer all-reduce unavailable (%s); using NCCL", e)
                ok = False
        flag = torch.tensor([1 if ok else 0], device=self.pack.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks or none (also: everyone's zeros are in place)
        torch.cuda.synchronize()
        if int(flag.item()) != 1:
            return False
        peers = nv.PinnPeers()
        peers.rank, peers.world, peers.epoch = rank, world, 0
        for r in range(world):
            peers.base[r] = int(hdl.buffer_ptrs[r])
        self._peers, self._peer_keep = peers, (region, hdl)
        return True

    def set_domains(self, domains: Sequence[FusedDomain]) -> None:
        """Swap in another binding of the same domains (fresh points); domain weights are re-read."""
        self.domains = list(domains)
        vals = [d.sigma for d in self.domains]
        if vals != self._sigma_vals:
            self._sigma_vals = vals
            self.sigmas = torch.tensor(vals, device=self.pack.device, dtype=torch.float32)

    @property
    def grad(self) -> torch.Tensor:
        return self.flat[: self.pack.n_params]

    @property
    def losses(self) -> torch.Tensor:
        """Per-domain un-weighted losses of the most recent launch."""
        return self._last[self.pack.n_params:]

    def launch(self, timed: Optional[Tuple[int, "torch.cuda.Event", "torch.cuda.Event"]] = None,
               loss_only: bool = False) -> None:
        """Enqueue the step kernels on the current stream (no host sync).
        `timed=(i, start, stop)` brackets domain i's kernel with CUDA events.  `loss_only` evaluates the
        losses without the reverse pass (pinn_loss_fused); `grad` is then unspecified."""
        lib = nv.lib()
        st = _stream_ptr()
        pack = self.pack
        pack.refresh_effective()
        mlp = pack.mlp_struct()
        for i, d in enumerate(self.domains):
            if timed is not None and timed[0] == i:
                timed[1].record()
            fn = lib.pinn_loss_fused if loss_only else lib.pinn_step_fused
            nv.check(fn(C.byref(mlp), C.byref(d.jets), C.byref(d.struct), self.ws.data_ptr(), self.ws_bytes, i, 1 if i else 0, st))
            if timed is not None and timed[0] == i:
                timed[2].record()
        if loss_only:
            if self._flat_eval is None:
                self._flat_eval = torch.zeros_like(self.flat)
            out = self._flat_eval
        else:
            out = self.flat
        self._last = out
        if self._peers is not None and not loss_only:
            self._peers.epoch += 1
            nv.check(lib.pinn_step_finalize_allreduce(C.byref(mlp), C.byref(self.domains[0].jets), self.ws.data_ptr(),
                                                      self.ws_bytes, len(self.domains), out.data_ptr(), C.byref(self._peers), st))
            self.reduced = True
            return
        self.reduced = False
        nv.check(lib.pinn_step_finalize(C.byref(mlp), C.byref(self.domains[0].jets), self.ws.data_ptr(), self.ws_bytes,
                                        len(self.domains), out.data_ptr(), 0, out.data_ptr() + 4 * pack.n_params, st))

    def total_loss(self) -> torch.Tensor:
        return torch.dot(self.losses, self.sigmas)

    def step(self, scatter: bool = True) -> torch.Tensor:
        """launch + (optional) write `.grad` of the module parameters; returns the
        total loss as a 0-d device tensor."""
        self.launch()
        if scatter:
            self.pack.scatter_grad(self.grad)
        return self.total_loss()
