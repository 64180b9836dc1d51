"""gnnx -- B200-native GNNExplainer mask-optimisation engine (drop-in for the hot path of
RexYing/gnn-model-explainer: Explainer.explain/explain_nodes, models.GraphConv/GcnEncoderNode,
graph_utils.neighborhoods).  All compute is in libgnnx.so (hand-written sm_100a CUDA behind a
C ABI, include/gnnx.h); there is no CPU fallback."""
from . import _abi  # noqa: F401
from .engine import Engine, Plan  # noqa: F401
from .explain import Explainer  # noqa: F401
from . import graph_utils, io_utils, models  # noqa: F401

__all__ = ["Engine", "Plan", "Explainer", "graph_utils", "io_utils", "models"]
