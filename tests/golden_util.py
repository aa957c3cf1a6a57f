"""Helpers for reading the golden fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["simple_poisson", "burgers", "allen_cahn", "ldc", "ns_unsteady", "ns_unsteady_wide"]
# the mlp_xl net at its true shape 3-512x6-3 (BASELINE configs[4]); gradients of the reference's FP64 run only,
# stored rounded to float32 (see make_golden.COMPACT)
XL = "ns_unsteady_xl"


def load(name):
    from oracle.pinn_oracle import problem_from_flat
    flat = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
    return problem_from_flat(flat), flat


def ref_domain(flat, tag, dom):
    pre = f"{tag}/dom/{dom}/"
    return {k[len(pre):]: v for k, v in flat.items() if k.startswith(pre)}


def ref_grads(flat, tag):
    pre = f"{tag}/grad/"
    return {k[len(pre):]: v for k, v in flat.items() if k.startswith(pre)}


def ref_components(flat, tag):
    return json.loads(str(flat[f"{tag}/loss_components"]))


def normwise(a, b):
    """max|a-b| / max|b|  (parity is norm-wise per tensor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)
