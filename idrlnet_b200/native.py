"""ctypes binding of libpinn_b200.so (include/pinn_b200.h).

The structures here mirror the header one to one.  `lib()` loads the in-tree
shared library and raises if it is missing: there is no CPU fallback -- a
product call without the CUDA library fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

MAX_LAYERS = 16
MAX_IN = 4
MAX_OUT = 4
MAX_FIRST = 4
MAX_AUX = 8
MAX_EQ = 8
MAX_TERMS = 64
MAX_FACTORS = 7
SYM_COORD0 = 28
SYM_AUX0 = 32
SYM_COUNT = 40

ACT = {"tanh": 0, "silu": 1, "swish": 1, "sin": 2}
LOSS = {"square": 0, "L1": 1, "Identity": 2}

E_INVALID, E_UNSUPPORTED, E_WORKSPACE, E_NO_DEVICE = -1, -2, -3, -4

_fp = C.c_void_p  # device (or, for the emulator, host) pointer to float


class PinnMlp(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("n_hidden", C.c_int32), ("width", C.c_int32),
                ("act", C.c_int32), ("precision", C.c_int32), ("W", _fp * MAX_LAYERS), ("b", _fp * MAX_LAYERS)]


class PinnJets(C.Structure):
    _fields_ = [("n_first", C.c_int32), ("n_second", C.c_int32), ("first_in", C.c_int32 * MAX_FIRST),
                ("n_mixed", C.c_int32), ("n_high", C.c_int32)]


class PinnTerm(C.Structure):
    _fields_ = [("coef", C.c_float), ("n_fac", C.c_uint8), ("fac", C.c_uint8 * MAX_FACTORS)]


class PinnEquation(C.Structure):
    _fields_ = [("term_begin", C.c_int32), ("term_end", C.c_int32), ("target", _fp), ("lambda_", _fp),
                ("target_const", C.c_float), ("lambda_const", C.c_float)]


class PinnDomain(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("coords", _fp * MAX_IN), ("aux", _fp * MAX_AUX), ("area", _fp),
                ("area_const", C.c_float), ("sigma", C.c_float), ("n_aux", C.c_int32), ("loss_kind", C.c_int32),
                ("n_eq", C.c_int32), ("n_terms", C.c_int32), ("eq", PinnEquation * MAX_EQ),
                ("terms", PinnTerm * MAX_TERMS)]


PRECISION = {"fp32": 0, "tf32": 1, "tf32_streamed": 2, "fp32x3": 3}


def make_mlp(n_in: int, n_out: int, n_hidden: int, width: int, act: str, w_ptrs: Sequence[int],
             b_ptrs: Sequence[int], precision: str = "fp32") -> PinnMlp:
    m = PinnMlp()
    m.n_in, m.n_out, m.n_hidden, m.width, m.act = n_in, n_out, n_hidden, width, ACT[act]
    m.precision = PRECISION[precision]
    assert len(w_ptrs) == n_hidden + 1 == len(b_ptrs)
    for i, (w, b) in enumerate(zip(w_ptrs, b_ptrs)):
        m.W[i], m.b[i] = w, b
    return m


def make_jets(first_in: Sequence[int], n_second: int, n_mixed: int = 0, n_high: int = 0) -> PinnJets:
    j = PinnJets()
    j.n_first, j.n_second, j.n_mixed, j.n_high = len(first_in), n_second, n_mixed, n_high
    for i, d in enumerate(first_in):
        j.first_in[i] = d
    return j


MAX_RANKS = 16


class PinnPeers(C.Structure):
    """pinn_peers (include/pinn_b200.h)."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("epoch", C.c_uint32), ("reserved", C.c_int32),
                ("base", C.c_void_p * MAX_RANKS)]


class PinnWnLayer(C.Structure):
    """pinn_wn_layer (include/pinn_b200.h)."""
    _fields_ = [("g", C.c_void_p), ("v", C.c_void_p), ("dW", C.c_void_p), ("W", C.c_void_p), ("dg", C.c_void_p),
                ("dv", C.c_void_p), ("out_dim", C.c_int32), ("in_dim", C.c_int32)]


class PinnShapeOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("p", C.c_float * 6)]


class PinnExprOp(C.Structure):
    """pinn_expr_op (include/pinn_b200.h)."""
    _fields_ = [("op", C.c_int32), ("imm", C.c_float)]


EXPR_MAX_OUT, EXPR_MAX_VARS, EXPR_MAX_OPS, EXPR_STACK = 12, 8, 448, 16
EX = {"const": 0, "var": 1, "add": 2, "sub": 3, "mul": 4, "div": 5, "min": 6, "max": 7, "pow": 8, "atan2": 9, "neg": 16, "abs": 17,
      "sqrt": 18, "sin": 19, "cos": 20, "tan": 21, "exp": 22, "log": 23, "tanh": 24, "sign": 25, "heaviside": 26, "square": 27,
      "rcp": 28, "scale": 29, "offset": 30, "asin": 31, "acos": 32, "atan": 33, "sinh": 34, "cosh": 35}

SHAPE_OPS = {"interval": 0, "rect": 1, "circle": 2, "channel2d": 3, "box": 4, "sphere": 5,
             "union": 16, "difference": 17, "intersection": 18, "complement": 19}


class PinnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpinn_b200 error {code}: {msg}")
        self.code = code


_LIB: Optional[C.CDLL] = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpinn_b200.so")

_SIGNATURES = {
    "pinn_abi_version": (C.c_int, []),
    "pinn_param_count": (C.c_int64, [C.POINTER(PinnMlp)]),
    "pinn_workspace_bytes": (C.c_int64, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_int32]),
    "pinn_supported": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets)]),
    "pinn_launch_count": (C.c_int64, []),
    "pinn_last_error": (C.c_char_p, []),
    "pinn_step_fused": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.POINTER(PinnDomain), C.c_void_p,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pinn_loss_fused": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.POINTER(PinnDomain), C.c_void_p,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pinn_step_finalize": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "pinn_allreduce_region_bytes": (C.c_int64, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_int32, C.c_int32]),
    "pinn_step_finalize_allreduce": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_void_p, C.c_int64, C.c_int32,
                                               C.c_void_p, C.POINTER(PinnPeers), C.c_void_p]),
    "pinn_jet_forward": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_int64, C.POINTER(_fp), C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "pinn_jet_backward": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_int64, C.POINTER(_fp), C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "pinn_weight_norm_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "pinn_weight_norm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_int32, C.c_void_p]),
    "pinn_weight_norm_forward_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pinn_weight_norm_backward_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pinn_sample_box": (C.c_int, [C.POINTER(_fp), C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64,
                                  C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p]),
    "pinn_residual_workspace_bytes": (C.c_int64, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.c_int64]),
    "pinn_residual_forward": (C.c_int, [C.POINTER(PinnMlp), C.POINTER(PinnJets), C.POINTER(PinnDomain), C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p]),
    "pinn_topk_workspace_bytes": (C.c_int64, [C.c_int32]),
    "pinn_topk_abs": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p]),
    "pinn_sample_csg_workspace_bytes": (C.c_int64, [C.c_int64]),
    "pinn_sample_csg": (C.c_int, [C.POINTER(_fp), C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, C.c_int64,
                                  C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(PinnShapeOp), C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "pinn_sample_program": (C.c_int, [C.POINTER(_fp), C.c_int32, C.POINTER(PinnExprOp), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "pinn_fp32_fma_probe": (C.c_int64, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pinn_fp32_fma2_probe": (C.c_int64, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pinn_tf32_mma_probe": (C.c_int64, [C.c_int32, C.c_void_p, C.c_void_p]),
    "pinn_tc_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """The loaded CUDA library.  Raises (never falls back) when it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C idrlnet_b200/csrc`).  idrlnet_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.pinn_abi_version() != 2:
            raise ImportError("libpinn_b200.so ABI version mismatch; rebuild")
        _LIB = handle
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        raise PinnError(rc, lib().pinn_last_error().decode(errors="replace"))


def ptr_array(ptrs: Sequence[int]):
    arr = (_fp * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
