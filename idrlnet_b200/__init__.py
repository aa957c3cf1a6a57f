"""idrlnet_b200 -- B200-native PINN training step behind IDRLnet's NetNode /
PdeNode / Solver API.  `import idrlnet_b200.shortcut as sc` replaces
`import idrlnet.shortcut as sc`.

Device selection mirrors idrlnet/__init__.py:20-43 (`use_gpu` / `use_cpu`), except
that a visible CUDA device is used by default: the fused path only exists on the
GPU (there is no CPU implementation of it; `use_cpu()` selects the generic
torch-autograd path, which is the reference's own algorithm)."""
import os

import torch

from .header import logger

GPU_AVAILABLE = torch.cuda.is_available()
GPU_ENABLED = False
DEVICE = torch.device("cpu")


def use_gpu(device=None):
    """Run on CUDA device `device` (default: LOCAL_RANK, else 0)."""
    global GPU_ENABLED, DEVICE
    if not GPU_AVAILABLE:
        logger.warning("use_gpu(): no CUDA device available")
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device(device if isinstance(device, (str, torch.device)) else f"cuda:{int(device)}")
    torch.cuda.set_device(dev)
    DEVICE, GPU_ENABLED = dev, True
    logger.info(f"Using GPU device {dev}")


def use_cpu():
    global GPU_ENABLED, DEVICE
    DEVICE, GPU_ENABLED = torch.device("cpu"), False
    logger.info("Using CPU (generic torch path; the fused CUDA step is disabled)")


if GPU_AVAILABLE and os.environ.get("IDRLNET_B200_DEVICE", "cuda") != "cpu":
    use_gpu()
