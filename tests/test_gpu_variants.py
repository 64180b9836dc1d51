"""Model variants of SURVEY 8 row f3 on the GPU (explain_var.cu through the C ABI): num_gc_layers = 2 / 4 and --bn against the
masks the UNMODIFIED reference returned (tests/golden/variants_golden.npz, oracle/gen_golden.py --only variants; 30 epochs),
plus random models (bn with 2 and 4 layers, odd widths) against the line-by-line torch port."""
import types

import numpy as np
import pytest

import gnnx
import gnnx_oracle as O
import util
from gnnx import _abi

pytestmark = pytest.mark.gpu


def _engine(rowptr, col, feat, label, pred_label, w, L, bn):
    eng = gnnx.Engine(0)
    eng.set_model(w, num_layers=L, bn=bn)
    eng.set_graph_csr(rowptr, col, feat, label, pred_label)
    return eng


@pytest.mark.parametrize("tag,L,bn", [("L2", 2, False), ("L4", 4, False), ("bn", 3, True)])
def test_variant_matches_reference_golden(tag, L, bn):
    g = np.load(util.GOLDEN + "/variants_golden.npz")
    N = int(g["N"]); epochs = int(g["num_epochs"])
    rowptr, col = O.csr_from_edges(N, g["edges"])
    w = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_W") or k.startswith(tag + "_b")}
    pred_label = np.argmax(g[tag + "_pred"], 1).astype(np.int32)
    nodes = [int(x) for x in g[tag + "_nodes"]]
    eng = _engine(rowptr, col, g["feat"].astype(np.float32), g["label"].astype(np.int32), pred_label, w, L, bn)
    plan = eng.plan_nodes(nodes, L)          # explain.py:64: n_hops = num_gc_layers
    m0 = np.empty(plan.total_edges, np.float32)
    for t, node in enumerate(nodes):
        assert np.array_equal(plan.neighbors_of(t), g["%s_n%d_nbrs" % (tag, node)]) and int(plan.node_idx_new[t]) == int(g["%s_n%d_idx_new" % (tag, node)])
        m0[plan.edge_off[t]:plan.edge_off[t + 1]] = g["%s_n%d_m0" % (tag, node)]
    out = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), m0, out)
    for t, node in enumerate(nodes):
        err = util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], g["%s_n%d_mask" % (tag, node)])
        assert err <= 1e-4, (tag, node, err)
    # num_epochs = 1 returns the initial mask, and the optimisation is deterministic
    one = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=1), m0, one)
    S = 1 / (1 + np.exp(-m0.astype(np.float64)))
    again = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=epochs), m0, again)
    assert np.array_equal(again, out) and np.isfinite(one).all() and abs(one.mean() - S.mean()) < 1e-6
    eng.close()


@pytest.mark.parametrize("seed,L,bn,hid,emb,d,C", [(1, 2, True, 20, 20, 10, 4), (2, 4, True, 20, 20, 7, 3), (3, 3, True, 16, 12, 33, 5),
                                                  (4, 2, False, 32, 8, 5, 2), (5, 4, False, 24, 24, 128, 6)])
def test_variant_matches_oracle_random(seed, L, bn, hid, emb, d, C):
    import networkx as nx
    import torch
    rng = np.random.default_rng(seed)
    N = 48
    G = nx.barabasi_albert_graph(N, 2, seed=seed)
    rowptr, col = O.csr_from_edges(N, np.array(G.edges(), dtype=np.int64))
    A = O.dense_from_csr(rowptr, col)
    feat = rng.normal(size=(N, d)).astype(np.float32)
    label = rng.integers(0, C, N).astype(np.int32)
    sc = lambda *s: (rng.normal(size=s) * 0.5).astype(np.float32)
    w = {}
    dims = [d] + [hid] * (L - 1) + [emb]
    for l in range(1, L + 1):
        w["W%d" % l] = sc(dims[l - 1], dims[l]); w["b%d" % l] = sc(dims[l])
    w["Wp"] = sc(C, hid * (L - 1) + emb); w["bp"] = sc(C)
    with torch.no_grad():
        pred = O._gcn_forward_torch(torch.tensor(feat[None]), torch.tensor(A[None], dtype=torch.float), O.weights_to_torch(w, False), False, bn=bn)[0].numpy()
    pred_label = np.argmax(pred, 1).astype(np.int32)
    nodes = [0, 7, 23, 47]
    eng = _engine(rowptr, col, feat, label, pred_label, w, L, bn)
    plan = eng.plan_nodes(nodes, L)
    m0 = np.empty(plan.total_edges, np.float32)
    dense_m0 = []
    for t in range(plan.count):
        M0 = O.draw_m0(plan.n(t), seed=500 * seed + t)
        r, c = plan.rows_cols_of(t)
        m0[plan.edge_off[t]:plan.edge_off[t + 1]] = M0[r, c]
        dense_m0.append(M0)
    out = np.zeros(plan.total_edges, np.float32)
    fm = np.zeros((plan.count, d), np.float32)
    E = 20
    eng.explain_nodes_host(eng.make_hparams(num_epochs=E), m0, out, fm)
    for t, node in enumerate(nodes):
        idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(rowptr, col, feat, label, node, L)
        Asub = O.dense_from_csr(srp, scol)
        hp = O.default_hparams(num_epochs=E)
        ref = O.explain_dense_torch(Asub, sfeat, slabel[idx], pred_label[nbrs], idx, w, dense_m0[t], hp=hp, bn=bn)
        c64 = O.explain_closed_form(Asub, sfeat, slabel[idx], pred_label[nbrs], idx, w, dense_m0[t], hp=hp, bn=bn)
        tol = max(1e-4, 3 * O.rel_l2(c64, ref))
        got = plan.dense_of(t, out)
        assert O.rel_l2(got, ref) <= tol, (node, O.rel_l2(got, ref), tol)
    assert np.isfinite(fm).all() and (fm > 0).all() and (fm < 1).all()
    eng.close()


def test_variant_refuses_what_it_does_not_build():
    g = np.load(util.GOLDEN + "/variants_golden.npz")
    N = int(g["N"])
    rowptr, col = O.csr_from_edges(N, g["edges"])
    w = {k[3:]: g[k] for k in g.files if k.startswith("bn_W") or k.startswith("bn_b")}
    eng = _engine(rowptr, col, g["feat"].astype(np.float32), g["label"].astype(np.int32), np.argmax(g["bn_pred"], 1).astype(np.int32), w, 3, True)
    plan = eng.plan_nodes([0], 3)
    out = np.zeros(plan.total_edges, np.float32)
    with pytest.raises(Exception):
        eng.grad_nodes_host(out)
    with pytest.raises(Exception):
        eng.explain_nodes_ex(eng.make_hparams(num_epochs=3), g["bn_n0_m0"].astype(np.float32), out, trace=np.zeros((1, 3, _abi.GX_TRACE_COLS), np.float32))
    eng.close()


OPT_CASES = [("sgd", dict(opt=1)), ("rmsprop", dict(opt=2)), ("adagrad", dict(opt=3)),
             ("adamstep", dict(opt=0, opt_scheduler=1, opt_decay_step=8, opt_decay_rate=0.5)),
             ("adamcos", dict(opt=0, opt_scheduler=2, opt_restart=12)),
             ("sgdstep", dict(opt=1, opt_scheduler=1, opt_decay_step=10, opt_decay_rate=0.3))]


@pytest.mark.parametrize("tag,over", OPT_CASES, ids=[c[0] for c in OPT_CASES])
@pytest.mark.parametrize("stream", [False, True], ids=["smem", "stream"])
def test_optimiser_variants_match_reference_golden(tag, over, stream):
    """utils/train_utils.py:7-23 variants (--opt sgd / rmsprop / adagrad, --opt-scheduler step / cos) against the masks the
    UNMODIFIED reference returned on the rand fixture (tests/golden/opts_golden.npz, oracle/gen_golden.py --only opts; 30 epochs).
    Adam + scheduler runs in the tuned kernels (shared-memory and streaming), the other optimisers in explain_var.cu."""
    if stream and over["opt"] != 0:
        pytest.skip("optimisers other than Adam always run in the variant kernel")
    g = np.load(util.GOLDEN + "/opts_golden.npz")
    fx = util.load_fixture("rand")
    eng = util.make_engine(fx)
    if stream:
        eng.debug_force_stream(True)
    plan = eng.plan_nodes(fx.nodes, 3)
    out = np.zeros(plan.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=int(g["num_epochs"]), **over), util.golden_m0(fx, plan), out)
    eng.close()
    for t, node in enumerate(fx.nodes):
        err = util.rel_l2(out[plan.edge_off[t]:plan.edge_off[t + 1]], g["%s_n%d_mask" % (tag, node)])
        assert err <= 1e-4, (tag, node, err)


def test_model_forward_matches_reference_predictions(tmp_path):
    """gx_model_forward (GcnEncoderNode.forward on the whole graph, models.py:58-80,230-267,363-376) against the logits the
    UNMODIFIED reference produced: syn1 / syn4 (trained checkpoints, golden/*_graph.npz `pred`) and the 2-layer, 4-layer and --bn
    models of variants_golden.npz; and Explainer(pred=None) uses it to obtain the predicted labels."""
    for name in ("syn1", "syn4", "rand"):
        fx = util.load_fixture(name)
        eng = util.make_engine(fx)
        got = eng.model_forward()
        eng.close()
        assert got.shape == fx.pred.shape
        assert np.abs(got - fx.pred).max() <= 2e-5 * max(1.0, np.abs(fx.pred).max()), name
        assert np.array_equal(np.argmax(got, 1), fx.pred_label)
    g = np.load(util.GOLDEN + "/variants_golden.npz")
    N = int(g["N"])
    rowptr, col = O.csr_from_edges(N, g["edges"])
    for tag, L, bn in (("L2", 2, False), ("L4", 4, False), ("bn", 3, True)):
        w = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_W") or k.startswith(tag + "_b")}
        ref = g[tag + "_pred"]
        eng = _engine(rowptr, col, g["feat"].astype(np.float32), g["label"].astype(np.int32), np.argmax(ref, 1).astype(np.int32), w, L, bn)
        got = eng.model_forward()
        eng.close()
        assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), tag
    # the drop-in without stored predictions
    import torch
    import test_gpu_parity as P
    fx = util.load_fixture("syn1")
    ex, args = P._explainer(fx, tmp_path, num_epochs=10)
    A = O.dense_from_csr(fx.rowptr, fx.col)
    nopred = gnnx.Explainer(model=ex.model, adj=A[None], feat=fx.feat[None].astype(np.float64), label=fx.label[None], pred=None,
                            train_idx=list(range(fx.N)), args=args, writer=None, print_training=False, graph_idx=-1)
    torch.manual_seed(3); a = nopred.explain(300)
    torch.manual_seed(3); b = ex.explain(300)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("seed,L,bn,hid,emb,d,C", [(11, 3, False, 64, 64, 10, 4), (12, 3, True, 48, 40, 17, 3), (13, 2, False, 128, 96, 6, 2), (14, 4, True, 64, 33, 128, 5)])
def test_wide_models_match_oracle(seed, L, bn, hid, emb, d, C):
    """--hidden-dim / --output-dim above 32 (explainer_main.py:51-56): the variant kernel with 2 / 4 lane chunks per row, against the
    line-by-line torch port."""
    test_variant_matches_oracle_random(seed, L, bn, hid, emb, d, C)
