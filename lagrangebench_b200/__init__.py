"""B200-native GNS rollout engine: a drop-in for lagrangebench's per-step hot path.

Mirrors the reference's public names for that path -- ``case_builder``
(``lagrangebench/case_setup``), ``models.GNS`` (``lagrangebench/models``), ``infer`` /
``eval_rollout`` (``lagrangebench/evaluate``) -- on top of hand-written sm_100a CUDA
kernels behind the C ABI of ``include/lb200.h``.
"""

from . import models  # noqa: F401
from . import data  # noqa: F401
from .case_setup import CaseSetupFn, PiecewiseForce, case_builder  # noqa: F401
from .data import H5Dataset  # noqa: F401
from .defaults import defaults  # noqa: F401
from .evaluate import MetricsComputer, RolloutEngine, averaged_metrics, eval_rollout, infer  # noqa: F401
from .models import GNS  # noqa: F401
from .strats import push_forward_build  # noqa: F401
from .utils import NodeType, get_kinematic_mask  # noqa: F401

__version__ = "0.2.0"
