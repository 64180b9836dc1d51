"""First-contact script for the GPU box (not a pytest): khop + explain on the golden fixtures."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import util
import gnnx

for name in ["rand", "syn4", "syn1"]:
    fx = util.load_fixture(name)
    eng = util.make_engine(fx)
    t0 = time.time()
    plan = eng.plan_nodes(fx.nodes, 3)
    t1 = time.time()
    bad = 0
    for t, node in enumerate(fx.nodes):
        if not np.array_equal(plan.neighbors_of(t), fx.gold["n%d_nbrs" % node]): bad += 1
        if int(plan.node_idx_new[t]) != int(fx.gold["n%d_idx_new" % node]): bad += 1
    print(name, "plan %.1f ms, khop mismatches: %d, total_n %d total_e %d" % ((t1 - t0) * 1e3, bad, plan.total_nodes, plan.total_edges))
    m0 = util.golden_m0(fx, plan)
    hp = eng.make_hparams()
    out = np.zeros(plan.total_edges, np.float32)
    fm = np.zeros((plan.count, fx.feat.shape[1]), np.float32)
    t0 = time.time()
    eng.explain_nodes_host(hp, m0, out, fm)
    t1 = time.time()
    errs = []
    for t, node in enumerate(fx.nodes):
        errs.append(util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], fx.gold["n%d_mask" % node]))
    errs = np.array(errs)
    print(name, "explain %.1f ms for %d nodes; rel-L2 max %.3e median %.3e; worst node %d" % (
        (t1 - t0) * 1e3, plan.count, errs.max(), np.median(errs), fx.nodes[int(errs.argmax())]))
    print("   first errs", errs[:8])
    eng.close()
