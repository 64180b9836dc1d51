"""GPU (-m gpu): the per-epoch trace, teacher forcing and checkpoint/resume, through the C ABI
(gx_explain_nodes_ex / gx_explain_graphs_ex / gx_offedge_regularisers).

  * trace      vs the numbers the UNMODIFIED reference prints with print_training=True (tests/golden/trace_golden.npz);
  * teacher    one Adam step from the reference's own optimiser state at t0 in {25, 50, 75} must reproduce its state at t0+1 to
               1e-5 -- also on the five chaotic syn1 nodes (0, 3, 33, 163, 293), where the 100-epoch mask cannot be matched to 1e-4
               by ANY reordering of the arithmetic (tests/golden/syn1_cond.npz): parity of the per-step arithmetic is what is checkable;
  * resume     E epochs == E1 epochs, state out, GX_INIT_STATE for the rest -- bit for bit."""
import numpy as np
import pytest

import gnnx
from gnnx import _abi
import gnnx_oracle as O
import util
from test_oracle_state import dense_m0

pytestmark = pytest.mark.gpu


def _engine(fx, stream):
    eng = util.make_engine(fx)
    if stream:
        eng.debug_force_stream(True)
    return eng


@pytest.mark.parametrize("stream", [False, True])
@pytest.mark.parametrize("which", ["syn1", "rand"])
def test_trace_matches_what_the_reference_prints(which, stream):
    tg = np.load(util.GOLDEN + "/trace_golden.npz")
    E = int(tg["num_epochs"])
    fx = util.load_fixture(which)
    nodes = [int(x) for x in tg[which + "_nodes"]]
    eng = _engine(fx, stream)
    plan = eng.plan_nodes(nodes, 3)
    C = fx.weights["Wp"].shape[0]
    dense = [dense_m0(plan.n(t), int(fx.gold["n%d_seed" % node])) for t, node in enumerate(nodes)]
    m0 = np.empty(plan.total_edges, np.float32)
    for t in range(plan.count):
        r, c = plan.rows_cols_of(t)
        m0[plan.edge_off[t]:plan.edge_off[t + 1]] = dense[t][r, c]
    hp = eng.make_hparams(num_epochs=E)
    out = np.zeros(plan.total_edges, np.float32)
    trace = np.zeros((plan.count, E, _abi.GX_TRACE_COLS), np.float32)
    pred = np.zeros((plan.count, E, C), np.float32)
    eng.explain_nodes_ex(hp, m0, out, trace=trace, trace_pred=pred)
    off = eng.offedge_regularisers(hp, np.concatenate([D.ravel() for D in dense]))
    plain = np.zeros_like(out)
    eng.explain_nodes_host(hp, m0, plain)
    assert np.array_equal(plain, out), "requesting a trace changed the masks"
    for t, node in enumerate(nodes):
        ref = tg["%s_n%d_trace" % (which, node)]
        n = plan.n(t)
        loss = trace[t, :, _abi.TR_LOSS_EDGES].astype(np.float64) + hp.coef_size * off[t, :, 0] + hp.coef_ent * off[t, :, 1] / (n * n)
        assert np.abs(loss / ref[:, 0] - 1).max() <= 1e-5, (node, loss[:3], ref[:3, 0])
        assert np.abs(trace[t, :, _abi.TR_DENSITY] - ref[:, 1]).max() <= 1e-5, node
        assert np.abs(pred[t] - ref[:, 2:]).max() <= 1e-5, node
        parts = trace[t, :, [_abi.TR_PRED, _abi.TR_SIZE, _abi.TR_ENT, _abi.TR_LAP, _abi.TR_FEAT]].sum(0)
        assert np.abs(parts / trace[t, :, _abi.TR_LOSS_EDGES] - 1).max() <= 1e-5
        assert np.abs(trace[t, :, _abi.TR_PGT] - pred[t, :, int(fx.label[node])]).max() <= 1e-6
    eng.close()


@pytest.mark.parametrize("stream", [False, True])
@pytest.mark.parametrize("which", ["syn1", "rand"])
def test_teacher_forced_step_matches_reference_state(which, stream):
    tg = np.load(util.GOLDEN + "/teacher_golden.npz")
    fx = util.load_fixture(which)
    nodes = [int(x) for x in tg[which + "_nodes"]]
    d = fx.feat.shape[1]
    eng = _engine(fx, stream)
    plan = eng.plan_nodes(nodes, 3)
    te = plan.total_edges
    worst = 0.0
    for t0 in [int(x) for x in tg["steps"]]:
        cat = lambda suffix: np.concatenate([tg["%s_n%d_t%d_%s" % (which, node, t0, suffix)] for node in nodes]).astype(np.float32)
        M, m, v = cat("M"), cat("m"), cat("v")
        feat = np.stack([tg["%s_n%d_t%d_feat" % (which, node, t0)] for node in nodes]).astype(np.float32)
        assert len(M) == te
        out = np.zeros(te, np.float32); sF = np.zeros((plan.count, d), np.float32)
        so = dict(M=np.zeros(te, np.float32), m=np.zeros(te, np.float32), v=np.zeros(te, np.float32), feat=np.zeros((plan.count, 3, d), np.float32))
        hp = eng.make_hparams(num_epochs=2, init=_abi.GX_INIT_STATE, start_step=t0)
        eng.explain_nodes_ex(hp, M, out, feat_mask_out=sF, state_in=dict(m=m, v=v, feat=feat), state_out=so)
        for t, node in enumerate(nodes):
            sl = slice(plan.edge_off[t], plan.edge_off[t + 1])
            key = "%s_n%d_t%d_" % (which, node, t0)
            e_mask = util.rel_l2(out[sl], tg[key + "mask_next"])
            # the step itself: (M_{t0+1} - M_{t0}) is what one epoch of kernel arithmetic produces
            e_step = util.rel_l2(so["M"][sl] - M[sl], tg[key + "M_next"] - M[sl])
            e_M = util.rel_l2(so["M"][sl], tg[key + "M_next"])
            assert e_mask <= 1e-5 and e_M <= 1e-5, (node, t0, e_mask, e_M)
            assert e_step <= 2e-3, (node, t0, e_step)
            assert np.abs(sF[t] - tg[key + "sF_next"]).max() <= 1e-5, (node, t0)
            worst = max(worst, e_mask)
    print("teacher-forced worst rel-L2 of the next mask: %.2e (%s, %s)" % (worst, which, "stream" if stream else "smem"))
    eng.close()


@pytest.mark.parametrize("stream", [False, True])
def test_resume_is_bit_identical(stream):
    fx = util.load_fixture("syn1")
    nodes = [0, 300, 683, 13, 450]
    d = fx.feat.shape[1]
    eng = _engine(fx, stream)
    plan = eng.plan_nodes(nodes, 3)
    te = plan.total_edges
    m0 = util.golden_m0(fx, plan)
    full = np.zeros(te, np.float32); fm_full = np.zeros((plan.count, d), np.float32)
    eng.explain_nodes_host(eng.make_hparams(num_epochs=30), m0, full, fm_full)
    so = dict(M=np.zeros(te, np.float32), m=np.zeros(te, np.float32), v=np.zeros(te, np.float32), feat=np.zeros((plan.count, 3, d), np.float32))
    part = np.zeros(te, np.float32)
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=12), m0, part, state_out=so)           # 11 updates
    rest = np.zeros(te, np.float32); fm = np.zeros((plan.count, d), np.float32)
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=19, init=_abi.GX_INIT_STATE, start_step=11), so["M"], rest, feat_mask_out=fm,
                         state_in=dict(m=so["m"], v=so["v"], feat=so["feat"]))                 # 18 more = 29 = 30 epochs
    assert np.array_equal(rest, full) and np.array_equal(fm, fm_full)
    # num_epochs = 1 from a state: nothing is updated, the state comes back unchanged
    so1 = dict(M=np.zeros(te, np.float32), m=np.zeros(te, np.float32), v=np.zeros(te, np.float32), feat=np.zeros((plan.count, 3, d), np.float32))
    one = np.zeros(te, np.float32)
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=1, init=_abi.GX_INIT_STATE, start_step=11), so["M"], one,
                         state_in=dict(m=so["m"], v=so["v"], feat=so["feat"]), state_out=so1)
    assert np.array_equal(one, part) and all(np.array_equal(so1[k], so[k]) for k in so)
    eng.close()


def test_graph_mode_trace_and_resume():
    g = np.load(util.GOLDEN + "/graphs_golden.npz")
    W = {k: g[k] for k in util.WKEYS}
    eng = gnnx.Engine(0)
    eng.set_model(W)
    eng.set_graph_batch(g["adj"], g["feat"], g["label"])
    gids = [0, 3, 7]
    eoff = eng.plan_graphs(gids)
    te = int(eoff[-1]); d = g["feat"].shape[2]; n = int(g["max_nodes"]); C = W["Wp"].shape[0]
    m0 = np.concatenate([g["g%d_m0" % i] for i in gids]).astype(np.float32)
    E = 10
    full = np.zeros(te, np.float32)
    eng.explain_graphs_host(eng.make_hparams(num_epochs=E), m0, full)
    trace = np.zeros((len(gids), E, _abi.GX_TRACE_COLS), np.float32); pred = np.zeros((len(gids), E, C), np.float32)
    out = np.zeros(te, np.float32)
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=E), m0, out, trace=trace, trace_pred=pred, graphs=True)
    assert np.array_equal(out, full)
    for t, gi in enumerate(gids):      # against the line-by-line port (bit-exact to the reference in graph mode, tests/test_oracle.py)
        A = g["adj"][gi].astype(np.float64)
        ei, ej = np.nonzero(A)
        M0 = np.ones((n, n), np.float32); M0[ei, ej] = g["g%d_m0" % gi]      # off-edge entries: any value, removed below
        tr = []
        O.explain_dense_torch(A, g["feat"][gi], int(g["label"][gi]), None, 0, W, M0, hp=O.default_hparams(num_epochs=E), graph_mode=True, trace=tr)
        for e in range(E):
            edges = tr[e]["pred_loss"] + tr[e]["size_edges"] + tr[e]["ent_edges"] + tr[e]["lap"] + tr[e]["feat_size"]
            assert abs(trace[t, e, _abi.TR_LOSS_EDGES] - edges) <= 1e-5 * abs(edges), (gi, e)
            assert abs(trace[t, e, _abi.TR_DENSITY] - tr[e]["density"]) <= 1e-5
            assert np.abs(pred[t, e] - tr[e]["pred"]).max() <= 1e-5
    so = dict(M=np.zeros(te, np.float32), m=np.zeros(te, np.float32), v=np.zeros(te, np.float32), feat=np.zeros((len(gids), 3, d), np.float32))
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=4), m0, out, state_out=so, graphs=True)
    rest = np.zeros(te, np.float32)
    eng.explain_nodes_ex(eng.make_hparams(num_epochs=E - 3, init=_abi.GX_INIT_STATE, start_step=3), so["M"], rest,
                         state_in=dict(m=so["m"], v=so["v"], feat=so["feat"]), graphs=True)
    assert np.array_equal(rest, full)
    eng.close()
