#!/usr/bin/env python
"""Per-node parity report of the CUDA kernels against the reference goldens (GPU box).

For every fixture (syn1 / syn4 / rand), horizon (10 / 30 / 100 epochs), kernel (shared-memory / streaming) and edge-phase
arithmetic (hardware approximations / IEEE): relative L2 of every golden node's mask vs the reference's, next to the per-node
tolerance the tests use (tests/util.py:node_tolerances: 1e-4 wherever the reference itself is reproducible under +-1 ulp input noise).
Writes gpurun_out/parity_report.json and prints a table; profiles/r02_parity_report.md is the committed summary."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gnn-model-explainer_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import util  # noqa: E402


def main():
    rep = {}
    for name in ("syn1", "syn4", "rand"):
        fx = util.load_fixture(name)
        tols = {ep: util.node_tolerances(name, ep) for ep in (10, 30, 100)}   # the tests' per-node rule (tests/util.py)
        gold = {10: np.load(util.GOLDEN + "/%s_golden_e10.npz" % name), 30: np.load(util.GOLDEN + "/%s_golden_e30.npz" % name), 100: fx.gold}
        for stream in (False, True):
            for ieee in (False, True):
                eng = util.make_engine(fx)
                eng.debug_ieee_edge(ieee)
                if stream:
                    eng.debug_force_stream(True)
                plan = eng.plan_nodes(fx.nodes, 3)
                m0 = util.golden_m0(fx, plan)
                for ep in (10, 30, 100):
                    out = np.zeros(plan.total_edges, np.float32)
                    eng.explain_nodes_host(eng.make_hparams(num_epochs=ep), m0, out)
                    errs = {node: util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], gold[ep]["n%d_mask" % node]) for t, node in enumerate(fx.nodes)}
                    vals = np.array(list(errs.values()))
                    tol = tols[ep]
                    over = {str(n): [e, tol[n]] for n, e in errs.items() if e > 1e-4}
                    viol = {str(n): [e, tol[n]] for n, e in errs.items() if e > tol[n]}
                    key = "%s/%s/%s/e%d" % (name, "stream" if stream else "smem", "ieee" if ieee else "fast", ep)
                    rep[key] = dict(nodes=len(vals), within_1e4=int((vals <= 1e-4).sum()), median=float(np.median(vals)), max=float(vals.max()),
                                    over_1e4=over, over_tolerance=viol)
                    print("%-28s %3d/%3d <= 1e-4  median %.2e  max %.2e  >1e-4: %s  >tol: %s" % (
                        key, rep[key]["within_1e4"], len(vals), rep[key]["median"], rep[key]["max"],
                        {k: "%.1e" % v[0] for k, v in over.items()}, list(viol)), flush=True)
                eng.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
