"""CPU: pins the trace and the teacher-forcing fixtures against the oracle.
  * tests/golden/trace_golden.npz  = loss / mask density / softmax row PRINTED by the unmodified reference (print_training=True);
    the line-by-line port must print the same numbers, and its decomposition (edge part + off-edge part) must add up.
  * tests/golden/teacher_golden.npz = optimiser state of the unmodified reference after t0 and t0+1 Adam steps; the closed form
    (the kernels' specification) advanced ONE step from the state at t0 must land on the state at t0+1 -- on the chaotic syn1
    nodes too, because one step cannot amplify anything."""
import numpy as np
import pytest
import torch

import gnnx_oracle as O
import util


def dense_m0(n, seed):
    torch.manual_seed(seed)
    std = torch.nn.init.calculate_gain("relu") * (2.0 / (n + n)) ** 0.5
    return torch.FloatTensor(n, n).normal_(1.0, std).numpy()


@pytest.mark.parametrize("which", ["syn1", "rand"])
def test_port_prints_what_the_reference_prints(which):
    tg = np.load(util.GOLDEN + "/trace_golden.npz")
    E = int(tg["num_epochs"])
    fx = util.load_fixture(which)
    for node in [int(x) for x in tg[which + "_nodes"]][:2]:
        idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(fx.rowptr, fx.col, fx.feat, fx.label, node, 3)
        A = O.dense_from_csr(srp, scol)
        M0 = dense_m0(len(nbrs), int(fx.gold["n%d_seed" % node]))
        tr = []
        O.explain_dense_torch(A, sfeat, slabel[idx], fx.pred_label[nbrs], idx, fx.weights, M0, hp=O.default_hparams(num_epochs=E), trace=tr)
        ref = tg["%s_n%d_trace" % (which, node)]
        assert len(tr) == E
        for e in range(E):
            assert abs(tr[e]["loss"] - ref[e, 0]) <= 2e-6 * abs(ref[e, 0]), (node, e)
            assert abs(tr[e]["density"] - ref[e, 1]) <= 1e-6, (node, e)
            assert np.abs(tr[e]["pred"] - ref[e, 2:]).max() <= 2e-7, (node, e)
            parts = tr[e]["pred_loss"] + tr[e]["size_edges"] + tr[e]["size_off"] + tr[e]["ent_edges"] + tr[e]["ent_off"] + tr[e]["lap"] + tr[e]["feat_size"]
            assert abs(parts - tr[e]["loss"]) <= 5e-6 * abs(tr[e]["loss"])


@pytest.mark.parametrize("which", ["syn1", "rand"])
def test_closed_form_one_step_from_reference_state(which):
    tg = np.load(util.GOLDEN + "/teacher_golden.npz")
    fx = util.load_fixture(which)
    for node in [int(x) for x in tg[which + "_nodes"]]:
        idx, srp, scol, sfeat, slabel, nbrs = O.extract_neighborhood(fx.rowptr, fx.col, fx.feat, fx.label, node, 3)
        n = len(nbrs)
        if n > 320:
            continue                      # dense fp64 closed form on the hub's 464-node subgraph: covered on the GPU only
        A = O.dense_from_csr(srp, scol)
        ei, ej = np.nonzero(A)

        def dense(v):
            D = np.zeros((n, n)); D[ei, ej] = v
            return D
        for t0 in [int(x) for x in tg["steps"]]:
            key = "%s_n%d_t%d_" % (which, node, t0)
            st = dict(m=dense(tg[key + "m"]), v=dense(tg[key + "v"]), feat=tg[key + "feat"], step=t0)
            out, state = O.explain_closed_form(A, sfeat, slabel[idx], fx.pred_label[nbrs], idx, fx.weights, dense(tg[key + "M"]),
                                               hp=O.default_hparams(num_epochs=1), return_state=True, init_state=st)
            assert O.rel_l2(state["M"][ei, ej], tg[key + "M_next"]) <= 2e-6, (node, t0)
            S = 1 / (1 + np.exp(-state["M"]))
            assert O.rel_l2(((S + S.T) / 2)[ei, ej], tg[key + "mask_next"]) <= 2e-6, (node, t0)
            assert np.abs(1 / (1 + np.exp(-state["F"])) - tg[key + "sF_next"]).max() <= 2e-6, (node, t0)
