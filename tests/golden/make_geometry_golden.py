"""Golden samples of the REFERENCE geometry module (idrlnet/geo_utils), seeded NumPy RNG.
Run in the build container only:   python tests/golden/make_geometry_golden.py
Writes tests/golden/geometry.npz with keys "<case>/<sample>/<column>"."""
import os
import sys

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _install_shims  # noqa: E402
from geometry_cases import CASES, SEED  # noqa: E402


def main():
    _install_shims()
    import idrlnet.shortcut as sc
    out = {}
    for name, fn in CASES.items():
        np.random.seed(SEED)
        try:
            result = fn(sc, sp)
        except Exception as e:  # e.g. Heart: the reference's Heaviside lowering breaks under SymPy 1.14
            print(name, "SKIPPED: the reference itself fails here:", type(e).__name__, e)
            continue
        for sample, cols in result.items():
            for k, v in cols.items():
                out[f"{name}/{sample}/{k}"] = np.asarray(v, dtype=np.float64)
        print(name, "ok")
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
