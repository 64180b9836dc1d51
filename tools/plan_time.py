#!/usr/bin/env python
"""Wall time of gx_plan_nodes (k-hop count + host classes + fill, synchronised) for the 700-node syn1 list, and of a full step
(plan + explain, the bench's `value` path).    python tools/plan_time.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-model-explainer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import gnnx  # noqa: E402
from gnnx import _abi  # noqa: E402

g = bench.load_syn1("syn1")
eng = gnnx.Engine(0)
eng.set_model(g["weights"]); eng.set_graph_csr(g["rowptr"], g["col"], g["feat"], g["label"], g["pred_label"])
nodes = np.arange(g["N"], dtype=np.int32)
eng.plan_nodes(nodes, 3, fetch=False)
out = torch.empty(eng._plan_sizes[2], dtype=torch.float32, device="cuda")
hp = eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=3)
plan_us, step_ms, kern_ms = [], [], []
for _ in range(5):
    eng.plan_nodes(nodes, 3, fetch=False); eng.sync()
for _ in range(40):
    t0 = time.perf_counter(); eng.plan_nodes(nodes, 3, fetch=False); eng.sync(); plan_us.append((time.perf_counter() - t0) * 1e6)
for _ in range(25):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.plan_nodes(nodes, 3, fetch=False); eng.explain_nodes_ptr(hp, _abi.GX_DEVICE, 0, out.data_ptr()); eng.sync()
    step_ms.append((time.perf_counter() - t0) * 1e3); kern_ms.append(eng.last_explain_ms())
print(json.dumps({"plan_us_med": float(np.median(plan_us)), "plan_us_min": float(min(plan_us)), "step_ms_med": float(np.median(step_ms)), "step_ms_min": float(min(step_ms)),
                  "kernel_ms_med": float(np.median(kern_ms)), "outside_kernel_ms_med": float(np.median(np.array(step_ms) - np.array(kern_ms)))}))
