"""Thread-block-cluster launch class of the shared-memory kernel (explain_node.cu, CS = 2 / 4 CTAs share one task through DSMEM):
an opt-in latency tool for small batches (gx_debug_set_cluster; profiles/r02_interim_notes.md).  Same arithmetic per row; only
the order in which the per-warp dL/dsF partials are summed differs from the single-CTA run, so results agree to round-off."""
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
