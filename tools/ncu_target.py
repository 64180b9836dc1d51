"""Small, fixed workloads for `ncu --set full` (one capture per kernel family; B200_PROFILING.md recipe):
    ncu --set full --clock-control none --import-source on -k regex:explain_node_kernel -c 6 -o gpurun_out/r02_node python tools/ncu_target.py syn1
    ncu --set full --clock-control none --import-source on -k regex:explain_gang_kernel -c 1 -o gpurun_out/r02_gang python tools/ncu_target.py c5 [N] [tasks] [epochs]
    ncu ... -k regex:explain_graph_kernel -c 1 ... python tools/ncu_target.py graphs"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gnn-model-explainer_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import bench  # noqa: E402
import gnnx  # noqa: E402
from gnnx import _abi  # noqa: E402


def main():
    import torch
    what = sys.argv[1] if len(sys.argv) > 1 else "syn1"
    if what in ("syn1", "syn4"):
        g = bench.load_syn1(what)
        eng = gnnx.Engine(0)
        eng.set_model(g["weights"])
        eng.set_graph_csr(g["rowptr"], g["col"], g["feat"], g["label"], g["pred_label"])
        eng.plan_nodes(np.arange(g["N"], dtype=np.int32), 3, fetch=False)
        out = torch.empty(eng._plan_sizes[2], dtype=torch.float32, device="cuda")
        eng.explain_nodes_ptr(eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=3), _abi.GX_DEVICE, 0, out.data_ptr())
        torch.cuda.synchronize()
        print(what, "ok", float(out.sum()), eng.last_explain_ms())
    elif what == "c5":
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
        K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
        EP = int(sys.argv[4]) if len(sys.argv) > 4 else 3
        d, C = 128, 4
        rng = np.random.default_rng(0)
        rowptr, col = bench.make_ba_csr(N, 32, 0)
        X = rng.normal(size=(N, d)).astype(np.float32)
        sc = lambda *s_: (rng.normal(size=s_) * 0.3).astype(np.float32)
        W = dict(W1=sc(d, 20), b1=sc(20), W2=sc(20, 20), b2=sc(20), W3=sc(20, 20), b3=sc(20), Wp=sc(C, 60), bp=sc(C))
        eng = gnnx.Engine(0)
        eng.set_model(W)
        eng.set_graph_csr(rowptr, col, X, rng.integers(0, C, N).astype(np.int32), rng.integers(0, C, N).astype(np.int32))
        nodes = np.random.default_rng(1).permutation(N)[:K].astype(np.int32)
        eng.plan_nodes(nodes, 3, fetch=False)
        out = torch.empty(eng._plan_sizes[2], dtype=torch.float32, device="cuda")
        eng.explain_nodes_ptr(eng.make_hparams(num_epochs=EP, init=_abi.GX_INIT_PHILOX, seed=3), _abi.GX_DEVICE, 0, out.data_ptr())
        torch.cuda.synchronize()
        print("c5 ok", float(out.sum()), eng.last_explain_ms(), "sum_n", eng._plan_sizes[1], "sum_E_d", eng._plan_sizes[2])
    elif what == "graphs":
        adj, feat, label, W = bench.make_graph_batch()
        eng = gnnx.Engine(0)
        eng.set_model(W)
        eng.set_graph_batch(adj, feat, label)
        eoff = eng.plan_graphs(np.arange(adj.shape[0], dtype=np.int32))
        out = np.zeros(int(eoff[-1]), np.float32)
        eng.explain_graphs_host(eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=3), None, out)
        print("graphs ok", float(out.sum()))


if __name__ == "__main__":
    main()
