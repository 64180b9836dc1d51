"""ref_harness.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the UNMODIFIED reference from /root/reference (read-only, present only in
the authoring container, NOT on the GPU box) with the import shims SURVEY.md
section 8c lists, so that its own code can be executed to
  * validate the CPU restatement in oracle/gnnx_oracle.py, and
  * generate the golden vectors committed under tests/golden/ (gen_golden.py).

Shims (all forced by the container's package set, none changes arithmetic):
  1. MagicMock modules for matplotlib*, seaborn, tensorboardX* (absent here;
     only used for plotting/logging: explainer/explain.py:10-20).
  2. networkx>=3 removed to_numpy_matrix/from_numpy_matrix (gengraph.py:83,
     utils/graph_utils.py:39): aliased onto the array versions.
  3. io_utils.log_graph -> no-op (called inside gengraph.gen_syn4, gengraph.py:255).
Nothing under /root/reference is copied or modified.
"""
import os
import sys
import types
import contextlib
import io
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("GNNX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "explainer"))


_loaded = {}


def load():
    """Return a namespace with the reference modules (explain, models, graph_utils,
    gengraph, featgen, train, io_utils)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in [
        "matplotlib", "matplotlib.colors", "matplotlib.pyplot", "matplotlib.figure",
        "matplotlib.backends", "matplotlib.backends.backend_agg", "matplotlib.cm",
        "seaborn", "tensorboardX", "tensorboardX.utils",
    ]:
        sys.modules.setdefault(name, MagicMock())
    import numpy as np
    import networkx as nx
    if not hasattr(nx, "to_numpy_matrix"):
        nx.to_numpy_matrix = lambda G, *a, **k: np.asmatrix(nx.to_numpy_array(G, *a, **k))
    if not hasattr(nx, "from_numpy_matrix"):
        nx.from_numpy_matrix = nx.from_numpy_array
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.io_utils as io_utils
    io_utils.log_graph = lambda *a, **k: None
    import utils.graph_utils as graph_utils
    import utils.featgen as featgen
    import utils.train_utils as train_utils
    import models
    import gengraph
    import explainer.explain as explain
    import train
    _loaded.update(dict(io_utils=io_utils, graph_utils=graph_utils, featgen=featgen,
                        train_utils=train_utils, models=models, gengraph=gengraph,
                        explain=explain, train=train))
    return types.SimpleNamespace(**_loaded)


def explainer_args(**over):
    """Namespace with the defaults of explainer_main.py:143-167."""
    d = dict(logdir="/tmp/gnnx_ref_log", ckptdir="ckpt", dataset="syn1", bmname=None,
             opt="adam", opt_scheduler="none", cuda="0", lr=0.1, clip=2.0, batch_size=20,
             num_epochs=100, hidden_dim=20, output_dim=20, num_gc_layers=3, dropout=0.0,
             method="base", name_suffix="", explainer_suffix="", align_steps=1000,
             explain_node=None, graph_idx=-1, mask_act="sigmoid", mask_bias=False,
             multigraph_class=-1, multinode_class=-1, gpu=False, bias=True, bn=False,
             graph_mode=False, writer=False)
    d.update(over)
    os.makedirs(d["logdir"], exist_ok=True)
    return types.SimpleNamespace(**d)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
