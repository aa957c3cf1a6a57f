"""Golden outputs of the REFERENCE integral operators (idrlnet/pde_op/operator.py:195-426).
Run in the build container only:   python tests/golden/make_operator_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _install_shims  # noqa: E402
from operator_cases import CASES  # noqa: E402


def main():
    _install_shims()
    import idrlnet.shortcut as sc
    out = {}
    for name, fn in CASES.items():
        for k, v in fn(sc).items():
            out[f"{name}/{k}"] = np.asarray(v)
        print(name, "ok")
    np.savez_compressed(os.path.join(HERE, "operators.npz"), **out)


if __name__ == "__main__":
    main()
