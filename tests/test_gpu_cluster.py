"""Thread-block-cluster launch class of the shared-memory kernel (explain_node.cu, CS = 2 / 4 CTAs share one task through DSMEM).
Off by default; in latency mode (gx_debug_set_cluster(h, 0, 0)) gx_plan_nodes picks it for the most expensive tasks of a batch that leaves
SMs idle; cluster sizes 2 / 4 force it for every task.  Same arithmetic per row; only the order in which the per-warp dL/dsF partials are summed differs from the
single-CTA run, so results agree to round-off (which is why the default is off: a task's masks then never depend on the batch)."""
import numpy as np
import pytest

import util
from gnnx import _abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cs", [2, 4])
@pytest.mark.parametrize("name,epochs,golden", [("rand", 30, "rand_golden_e30.npz"), ("syn1", 10, "syn1_golden_e10.npz"), ("syn1", 100, "syn1_golden.npz")])
def test_cluster_matches_reference_golden(name, epochs, golden, cs):
    fx = util.load_fixture(name)
    g = np.load(util.GOLDEN + "/" + golden)
    nodes = fx.nodes if name != "syn1" else fx.nodes[:24]
    eng = util.make_engine(fx)
    eng.debug_cluster(cs, 1)             # every shared-memory task goes to the cluster class
    plan = eng.plan_nodes(nodes, 3)
    out = np.zeros(plan.total_edges, np.float32)
    fm = np.zeros((plan.count, fx.feat.shape[1]), np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), util.golden_m0(fx, plan), out, fm)
    again = np.zeros_like(out)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), util.golden_m0(fx, plan), again)
    eng.debug_cluster(1, 0)
    plan1 = eng.plan_nodes(nodes, 3)
    one = np.zeros(plan1.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=min(epochs, 10)), util.golden_m0(fx, plan1), one)
    eng.debug_cluster(cs, 1)
    eng.plan_nodes(nodes, 3)
    ten = np.zeros_like(out)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=min(epochs, 10)), util.golden_m0(fx, plan), ten)
    eng.close()
    assert np.array_equal(again, out)                       # deterministic
    assert util.rel_l2(ten, one) < 2e-6                     # 10 epochs: cluster == single CTA up to round-off
    errs = {node: util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], g["n%d_mask" % node]) for t, node in enumerate(nodes)}
    tol = util.node_tolerances(name, epochs)
    bad = {n: (e, tol[n]) for n, e in errs.items() if not e <= tol[n]}
    assert not bad, bad
    assert np.isfinite(fm).all()


def test_cluster_trace_and_philox():
    """The trace (a12) and the device Philox init under a cluster: same numbers as the single-CTA kernel up to round-off."""
    fx = util.load_fixture("rand")
    res = {}
    for cs in (1, 4):
        eng = util.make_engine(fx)
        eng.debug_cluster(cs, 1)
        plan = eng.plan_nodes(fx.nodes, 3)
        out = np.zeros(plan.total_edges, np.float32)
        tr = np.zeros((plan.count, 8, _abi.GX_TRACE_COLS), np.float32)
        eng.explain_nodes_ex(eng.make_hparams(num_epochs=8, init=_abi.GX_INIT_PHILOX, seed=5), None, out, trace=tr)
        eng.close()
        res[cs] = (out, tr)
    assert util.rel_l2(res[4][0], res[1][0]) < 2e-6
    assert np.allclose(res[4][1], res[1][1], rtol=2e-5, atol=1e-6)


def test_automatic_cluster_policy():
    """A small batch (the 24 most expensive syn1 nodes) gets clusters for its expensive tasks and agrees with the strict single-CTA run to
    round-off; the full 700-node batch has no spare SM and gets none (so its masks are the strict ones, bit for bit)."""
    fx = util.load_fixture("syn1")
    N = fx.rowptr.shape[0] - 1
    nodes = np.arange(24, dtype=np.int32)
    hp = dict(num_epochs=10, init=_abi.GX_INIT_PHILOX, seed=11)
    eng = util.make_engine(fx)
    plan = eng.plan_nodes(nodes, 3)
    assert eng.plan_class_counts()[0][6] == 0                     # default: no cluster class
    eng.debug_cluster(0, 0)                                       # latency mode
    plan = eng.plan_nodes(nodes, 3)
    counts, cs = eng.plan_class_counts()
    assert counts[6] > 0 and cs in (2, 4), (counts, cs)          # the automatic policy used the cluster class
    auto = np.zeros(plan.total_edges, np.float32); fa = np.zeros((plan.count, fx.feat.shape[1]), np.float32)
    eng.explain_nodes_host(eng.make_hparams(**hp), None, auto, fa)
    b, e = eng.last_class_ms()
    assert e[6] > 0 and (b[:6] < 0).sum() >= 3                    # timeline: the cluster class ran
    eng.debug_cluster(1, 0)
    plan1 = eng.plan_nodes(nodes, 3)
    counts1, cs1 = eng.plan_class_counts()
    assert counts1[6] == 0 and cs1 == 1
    off = np.zeros_like(auto); fo = np.zeros_like(fa)
    eng.explain_nodes_host(eng.make_hparams(**hp), None, off, fo)
    for t in range(len(nodes)):
        sl = slice(plan.edge_off[t], plan.edge_off[t + 1])
        assert util.rel_l2(auto[sl], off[sl]) < 2e-6, t
    assert np.allclose(fa, fo, rtol=1e-5, atol=1e-7)
    full = eng.plan_nodes(np.arange(N, dtype=np.int32), 3)
    countsf, csf = eng.plan_class_counts()
    assert countsf[6] == 0, countsf                               # strict plan of the full batch
    whole = np.zeros(full.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(**hp), None, whole)   # Philox streams are keyed by node id: same M0 as above
    eng.debug_cluster(0, 0)
    fulla = eng.plan_nodes(np.arange(N, dtype=np.int32), 3)
    countsa, _ = eng.plan_class_counts()
    assert countsa[6] == 0, countsa                               # automatic: a full batch has no spare SM
    wholea = np.zeros(fulla.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(**hp), None, wholea)
    eng.close()
    assert np.array_equal(whole, wholea)
    for t in range(len(nodes)):
        assert np.array_equal(whole[full.edge_off[t]:full.edge_off[t + 1]], off[plan.edge_off[t]:plan.edge_off[t + 1]]), t   # strict = batch independent
