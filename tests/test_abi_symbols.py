"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes as C
import os
import re

from idrlnet_b200 import native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pinn_b200.h")).read()
    return sorted(set(re.findall(r"PINN_API\s+[\w\s\*]+?\b(pinn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 14
    lib = C.CDLL(nv.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pinn_b200.h but not exported"
    assert sorted(nv.EXPORTS) == names  # the ctypes binding covers the whole header


def test_host_only_queries():
    lib = nv.lib()
    assert lib.pinn_abi_version() == 2
    mlp = nv.make_mlp(2, 1, 4, 20, "silu", [1] * 5, [1] * 5)  # pointers are not dereferenced by the queries
    assert lib.pinn_param_count(C.byref(mlp)) == 2 * 20 + 20 + 3 * (400 + 20) + 20 + 1 == 1341 + 0 * 1
    jets = nv.make_jets((0, 1), 1)
    assert lib.pinn_supported(C.byref(mlp), C.byref(jets)) == 1
    assert lib.pinn_workspace_bytes(C.byref(mlp), C.byref(jets), 3) > 0
    wide = nv.make_mlp(2, 3, 4, 100, "tanh", [1] * 5, [1] * 5)
    assert lib.pinn_supported(C.byref(wide), C.byref(nv.make_jets((0, 1), 2))) == 1
    assert lib.pinn_supported(C.byref(nv.make_mlp(2, 3, 4, 1000, "tanh", [1] * 5, [1] * 5)), C.byref(jets)) == 0
    assert lib.pinn_launch_count() == 0


def test_struct_sizes_match_header():
    # sizes implied by include/pinn_b200.h on LP64
    assert C.sizeof(nv.PinnTerm) == 12
    assert C.sizeof(nv.PinnEquation) == 32
    assert C.sizeof(nv.PinnJets) == 32
    assert C.sizeof(nv.PinnMlp) == 24 + 2 * 16 * 8
    assert C.sizeof(nv.PinnDomain) == 8 + 4 * 8 + 8 * 8 + 8 + 6 * 4 + 8 * 32 + 64 * 12
