"""API facade -- `import idrlnet_b200.shortcut as sc` replaces
`import idrlnet.shortcut as sc` (idrlnet/shortcut.py:1-13)."""
from .geo_utils import *  # noqa: F401,F403
from .architecture import *  # noqa: F401,F403
from .pde_op import *  # noqa: F401,F403
from .net import NetNode  # noqa: F401
from .data import get_data_node, DataNode, get_data_nodes, datanode, SampleDomain  # noqa: F401
from .pde import ExpressionNode, PdeNode  # noqa: F401
from .solver import Solver  # noqa: F401
from .callbacks import GradientReceiver, SummaryReceiver, HandleResultReceiver  # noqa: F401
from .receivers import Receiver, Signal  # noqa: F401
from .variable import Variables, export_var, Loss  # noqa: F401
from .header import logger  # noqa: F401
from . import GPU_AVAILABLE, GPU_ENABLED, use_gpu, use_cpu  # noqa: F401
