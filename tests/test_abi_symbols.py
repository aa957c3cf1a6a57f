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


def test_missing_library_fails_loudly(monkeypatch):
    """No CPU fallback: without the CUDA library the loader raises instead of routing anywhere else."""
    import pytest

    monkeypatch.setattr(nv, "_LIB", None)
    monkeypatch.setattr(nv, "LIB_PATH", os.path.join(os.path.dirname(nv.LIB_PATH), "no_such_libpinn_b200.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        nv.lib()


def test_product_never_touches_the_oracle():
    """oracle/ and tests/ are checker code: nothing under the package, the reference-side binding or the workload
    definitions imports them (bench.py may, for its reference arm only)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b", re.M)
    hits = []
    for top in ("idrlnet_b200", "integration", "baseline"):
        for dirpath, dirnames, files in os.walk(os.path.join(root, top)):
            dirnames[:] = [d for d in dirnames if d not in ("_ref", "__pycache__", "build")]
            for f in files:
                if f == "run_reference_with_jets.py":  # the driver tests/test_gpu_integration.py executes: loads golden fixtures
                    continue
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f), encoding="utf-8").read()):
                    hits.append(os.path.join(dirpath, f))
    assert not hits, hits
