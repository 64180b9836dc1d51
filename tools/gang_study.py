"""Streaming-kernel study on a config-5 style graph (GPU box): per-phase cycle counters (debug hook gx_debug_set_dump), kernel time
and masks of explain_gang.cu for several gang sizes against the first-generation kernel explain_stream.cu.
Usage: python tools/gang_study.py [N] [nodes] [epochs] [gang sizes, comma separated; -1 = explain_stream.cu]"""
import sys, os, json, numpy as np, ctypes as C, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gnn-model-explainer_b200'))
import bench, gnnx
from gnnx import _abi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
EP = int(sys.argv[3]) if len(sys.argv) > 3 else 20
GS = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [-1, 0, 1, 8, 37, 74, 148]
d, Cc = 128, 4
rng = np.random.default_rng(0)
rowptr, col = bench.make_ba_csr(N, 32, 0)
X = rng.normal(size=(N, d)).astype(np.float32)
sc = lambda *s_: (rng.normal(size=s_) * 0.3).astype(np.float32)
W = dict(W1=sc(d, 20), b1=sc(20), W2=sc(20, 20), b2=sc(20), W3=sc(20, 20), b3=sc(20), Wp=sc(Cc, 60), bp=sc(Cc))
eng = gnnx.Engine(0); eng.set_model(W)
eng.set_graph_csr(rowptr, col, X, rng.integers(0, Cc, N).astype(np.int32), rng.integers(0, Cc, N).astype(np.int32))
dbg = torch.zeros((1 << 19) + 64, dtype=torch.float32, device='cuda')
lib = _abi.lib()
lib.gx_debug_set_dump(eng._h, C.c_void_p(dbg.data_ptr()))
nodes = np.random.default_rng(1).permutation(N)[:K].astype(np.int32)
eng.plan_nodes(nodes, 3, fetch=False)
te = eng._plan_sizes[2]
hp = eng.make_hparams(num_epochs=EP, init=_abi.GX_INIT_PHILOX, seed=1)
ref = None
rows = []
for g in GS:
    eng.debug_gang(g)
    out = torch.zeros(te, dtype=torch.float32, device='cuda')
    dbg.zero_()
    ms = []
    for rep in range(2):
        eng.explain_nodes_ptr(hp, _abi.GX_DEVICE, 0, out.data_ptr()); torch.cuda.synchronize()
        ms.append(eng.last_explain_ms())
    D = dbg.cpu().numpy()[(1 << 19):(1 << 19) + 20]
    if g == -1:
        names = ['F0', 'F1', 'F2', 'S', 'B2', 'B1', 'B0', 'P']; ph = D[:8] / max(EP - 1, 1); meta = D[8:14]
    else:
        names = ['F0', 'F1', 'F2', 'S', 'B2', 'B1', 'B0s', 'B0d', 'P']; ph = D[:9] / max(EP - 1, 1); meta = D[9:17]
    o = out.double()
    if ref is None:
        ref = o.clone()
    rel = float(((o - ref).norm() / ref.norm()).item())
    row = dict(N=N, tasks=K, epochs=EP, gang=g, kernel_ms=min(ms), meta=[int(x) for x in meta], kcycles_per_epoch={nm: round(float(v) / 1e3, 1) for nm, v in zip(names, ph)},
               total_kcycles=round(float(ph.sum()) / 1e3, 1), rel_l2_vs_first=rel, finite=bool(torch.isfinite(o).all().item()))
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'gang_study_N%d_K%d.json' % (N, K)), 'w'), indent=1)
