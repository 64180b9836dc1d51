"""Streaming kernel (explain_stream.cu: tasks whose state does not fit shared memory, A_m (X W1) order) against the
reference goldens and the oracle, through the C ABI.  The small cases are forced into the streaming class with the
debug knob; the last test uses a graph whose 3-hop neighbourhoods are too large for shared memory on their own."""
import numpy as np
import pytest
import torch

import gnnx_oracle as O
import util
from test_gpu_parity import _random_case

pytestmark = pytest.mark.gpu


def _stream_engine(fx):
    eng = util.make_engine(fx)
    eng.debug_force_stream(True)
    return eng


@pytest.mark.parametrize("gang", [0, -1], ids=["gang", "stream1"])
@pytest.mark.parametrize("name,epochs,golden", [("syn1", 10, "syn1_golden_e10.npz"), ("syn1", 30, "syn1_golden_e30.npz"), ("syn1", 100, "syn1_golden.npz"),
                                                 ("syn4", 30, "syn4_golden_e30.npz"), ("syn4", 100, "syn4_golden.npz"), ("rand", 30, "rand_golden_e30.npz")])
def test_stream_matches_reference_golden(name, epochs, golden, gang):
    """Both streaming kernels (explain_gang.cu and the first-generation explain_stream.cu) against the reference goldens, per node
    (util.node_tolerances: 1e-4 wherever the reference itself is reproducible)."""
    fx = util.load_fixture(name)
    g = np.load(util.GOLDEN + "/" + golden)
    eng = _stream_engine(fx)
    eng.debug_gang(gang)
    plan = eng.plan_nodes(fx.nodes, 3)
    out = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), util.golden_m0(fx, plan), out)
    eng.close()
    errs = {node: util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], g["n%d_mask" % node]) for t, node in enumerate(fx.nodes)}
    util.assert_per_node(errs, name, epochs)
    assert np.median(list(errs.values())) < 2e-6


def test_gang_size_does_not_change_a_bit():
    """explain_gang.cu: the number of CTAs that share a task is a scheduling decision -- masks and feature masks are
    bit-identical for 1, 3, 16 CTAs per task and the automatic choice (what keeps sharded multi-GPU runs bit-identical)."""
    fx = util.load_fixture("rand")
    res = []
    for gang in (1, 3, 16, 0):
        eng = _stream_engine(fx)
        eng.debug_gang(gang)
        plan = eng.plan_nodes(fx.nodes, 3)
        out = np.zeros(plan.total_edges, np.float32)
        fm = np.zeros((plan.count, fx.feat.shape[1]), np.float32)
        eng.explain_nodes_host(eng.make_hparams(num_epochs=12), util.golden_m0(fx, plan), out, fm)
        eng.close()
        res.append((out, fm))
    for out, fm in res[1:]:
        assert np.array_equal(out, res[0][0]) and np.array_equal(fm, res[0][1])


@pytest.mark.parametrize("seed,n_nodes,m,d,C,graph", [
    (1, 40, 2, 10, 4, "ba"), (2, 30, 1, 1, 2, "ba"), (5, 9, 0, 5, 2, "star"), (6, 6, 0, 10, 4, "complete"),
    (7, 45, 0, 33, 5, "gnp"), (8, 35, 2, 64, 3, "ba"), (9, 30, 2, 128, 2, "ba"), (10, 50, 4, 32, 40, "ba"),
])
def test_stream_matches_oracle_random(seed, n_nodes, m, d, C, graph):
    cs = _random_case(seed, n_nodes, m, d, C, graph)
    eng = _stream_engine(cs)
    nodes = list(range(0, cs.N, max(1, cs.N // 5)))[:5]
    plan = eng.plan_nodes(nodes, 3)
    m0 = np.empty(plan.total_edges, np.float32)
    dense_m0 = []
    for t in range(plan.count):
        M0 = O.draw_m0(plan.n(t), seed=100 * seed + t)
        r, c = plan.rows_cols_of(t)
        m0[plan.edge_off[t]:plan.edge_off[t + 1]] = M0[r, c]
        dense_m0.append(M0)
    out = np.zeros(plan.total_edges, np.float32)
    fm = np.zeros((plan.count, d), np.float32)
    hp = eng.make_hparams(num_epochs=30)
    eng.explain_nodes_host(hp, m0, out, fm)
    for t, node in enumerate(nodes):
        idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(cs.rowptr, cs.col, cs.feat, cs.label, node, 3)
        A = O.dense_from_csr(srp, scol)
        ref = O.explain_dense_torch(A, sfeat, slabel[idx], cs.pred_label[nbrs], idx, cs.weights, dense_m0[t], hp=O.default_hparams(num_epochs=30))
        c64 = O.explain_closed_form(A, sfeat, slabel[idx], cs.pred_label[nbrs], idx, cs.weights, dense_m0[t], hp=O.default_hparams(num_epochs=30))
        tol = max(1e-4, 3 * O.rel_l2(c64, ref))
        got = plan.dense_of(t, out)
        assert O.rel_l2(got, ref) <= tol, (node, O.rel_l2(got, ref), tol)
        _, st = O.explain_closed_form(A, sfeat, slabel[idx], cs.pred_label[nbrs], idx, cs.weights, dense_m0[t],
                                      hp=O.default_hparams(num_epochs=29), return_state=True)
        assert np.abs(fm[t] - 1 / (1 + np.exp(-st["F"]))).max() < max(2e-4, 30 * O.rel_l2(c64, ref)), node
    eng.close()


def test_stream_one_epoch_and_philox_agree_with_resident_kernel():
    """num_epochs=1 returns the initial mask; with the device Philox init the two kernels must draw the same M0
    (they index the generator by (node, canonical edge slot)) and agree after 10 epochs."""
    fx = util.load_fixture("syn1")
    nodes = fx.nodes[:6]
    res = {}
    for stream in (False, True):
        eng = util.make_engine(fx)
        if stream:
            eng.debug_force_stream(True)
        plan = eng.plan_nodes(nodes, 3)
        for ep in (1, 10):
            out = np.zeros(plan.total_edges, np.float32)
            eng.explain_nodes_host(eng.make_hparams(num_epochs=ep, init=1, seed=7), None, out)
            res[(stream, ep)] = out
        eng.close()
    assert np.array_equal(res[(False, 1)], res[(True, 1)])
    assert util.rel_l2(res[(True, 10)], res[(False, 10)]) < 1e-5


@pytest.mark.parametrize("N,mm,d,nodes,epochs,min_n", [(1500, 6, 128, [0, 700, 1499], 10, 400), (5000, 4, 16, [10], 5, 4097)])
def test_stream_natural_class(N, mm, d, nodes, epochs, min_n):
    """BA graphs whose 3-hop neighbourhoods are most of the graph: the per-node state exceeds shared memory, so the plan
    puts the tasks into the streaming class by itself (second case: n > 4096, the plan's scan-based level ordering).
    Few epochs against the dense torch port."""
    import networkx as nx
    rng = np.random.default_rng(11)
    C = 4
    G = nx.barabasi_albert_graph(N, mm, seed=11)
    rowptr, col = O.csr_from_edges(N, np.array(G.edges(), dtype=np.int64))
    feat = rng.normal(size=(N, d)).astype(np.float32)
    label = rng.integers(0, C, N)
    sc = lambda *s: (rng.normal(size=s) * 0.3).astype(np.float32)
    w = dict(W1=sc(d, 20), b1=sc(20), W2=sc(20, 20), b2=sc(20), W3=sc(20, 20), b3=sc(20), Wp=sc(C, 60), bp=sc(C))
    A = np.zeros((N, N), np.float32)
    for i in range(N):
        A[i, col[rowptr[i]:rowptr[i + 1]]] = 1
    with torch.no_grad():
        pred = O._gcn_forward_torch(torch.tensor(feat[None]), torch.tensor(A[None]), O.weights_to_torch(w, False), False)[0].numpy()
    import types
    cs = types.SimpleNamespace(N=N, rowptr=rowptr, col=col, feat=feat, label=label, weights=w, pred_label=np.argmax(pred, 1).astype(np.int32))
    eng = util.make_engine(cs)
    plan = eng.plan_nodes(nodes, 3)
    assert min(plan.n(t) for t in range(plan.count)) >= min_n
    m0 = np.empty(plan.total_edges, np.float32)
    dense_m0 = []
    for t in range(plan.count):
        M0 = O.draw_m0(plan.n(t), seed=900 + t)
        r, c = plan.rows_cols_of(t)
        m0[plan.edge_off[t]:plan.edge_off[t + 1]] = M0[r, c]
        dense_m0.append(M0)
    out = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), m0, out)
    for t, node in enumerate(nodes):
        idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(rowptr, col, feat, label, node, 3)
        assert np.array_equal(nbrs, plan.neighbors_of(t))
        ref = O.explain_dense_torch(O.dense_from_csr(srp, scol), sfeat, slabel[idx], cs.pred_label[nbrs], idx, w, dense_m0[t],
                                    hp=O.default_hparams(num_epochs=epochs))
        r, c = plan.rows_cols_of(t)
        assert util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], ref[r, c]) <= 1e-4, node
    eng.close()
