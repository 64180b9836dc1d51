#!/usr/bin/env python
"""Cluster launch class study (GPU box): for cluster size / cost-threshold settings, parity of the golden nodes against the
reference masks and against the single-CTA result, and the device time of the 700-node syn1 batch (and syn4).
    python tools/cluster_study.py [syn1|syn4] """
import os
import sys
import json

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gnn-model-explainer_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import util  # noqa: E402
from gnnx import _abi  # noqa: E402


def run(name, cs, cost, reps=6, top=0):
    fx = util.load_fixture(name)
    eng = util.make_engine(fx)
    eng.debug_cluster(cs, cost)
    plan = eng.plan_nodes(fx.nodes, 3)
    out = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(), util.golden_m0(fx, plan), out)
    errs = np.array([util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], fx.gold["n%d_mask" % node]) for t, node in enumerate(fx.nodes)])
    N = fx.rowptr.shape[0] - 1
    allnodes = np.arange(N, dtype=np.int32) if top == 0 else np.arange(top, dtype=np.int32)   # syn1: the first nodes are the hubs = the most expensive tasks
    hp = eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=7)
    ms = []
    eng.plan_nodes(allnodes, 3, fetch=False)
    counts, plan_cs, smem = eng.plan_class_counts(with_smem=True)
    import torch
    te = eng._plan_sizes[2]
    dev_out = torch.empty(te, dtype=torch.float32, device="cuda")
    for _ in range(reps):
        eng.explain_nodes_device(hp, None, dev_out)
        eng.sync()
        ms.append(eng.last_explain_ms())
    cb, ce = eng.last_class_ms()
    res = dict(fixture=name, batch=len(allnodes), cluster=cs, cost=cost, kernel_ms_min=min(ms), kernel_ms_med=float(np.median(ms)), within_1e4=int((errs <= 1e-4).sum()), nodes=len(errs),
               max_err=float(errs.max()), checksum=float(dev_out.double().sum().item()),
               class_counts=[int(c) for c in counts], plan_cluster=plan_cs, class_smem_kb=[round(int(b) / 1024, 1) for b in smem],
               class_begin_ms=[round(float(x), 3) for x in cb], class_end_ms=[round(float(x), 3) for x in ce])
    eng.close()
    return res, out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "syn1"
    tops = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 24, 88, 175, 350, 0]
    modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 0, 2, 4]
    rows = []
    for top in tops:      # a single explain() call ... a shard of a strong-scaled run ... the full batch (0)
        ref = None
        for cs in modes:                         # 1 = clusters off, 0 = automatic policy, 2 / 4 = forced for every task above 200 000 cost units
            if top == 0 and cs in (2, 4):
                continue
            r, o = run(name, cs, 200000, top=top)
            if ref is None:
                ref = r["checksum"]
            r["checksum_equals_off"] = bool(r["checksum"] == ref)
            r["tag"] = os.environ.get("GNNX_STUDY_TAG", "")
            print(json.dumps(r), flush=True)
            rows.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "cluster_study_%s%s.json" % (name, os.environ.get("GNNX_STUDY_TAG", ""))), "w"), indent=1)


if __name__ == "__main__":
    main()
