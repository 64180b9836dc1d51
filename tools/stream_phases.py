"""Per-phase cycle counters of the streaming kernel (debug hook gx_debug_set_dump) on a config-5 style graph.
Usage on the GPU box: [GNNX_STREAM_THREADS=1024] python tools/stream_phases.py [N] [nodes]"""
import sys, os, numpy as np, ctypes as C, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gnn-model-explainer_b200'))
import bench, gnnx
from gnnx import _abi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
EP = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d, Cc = 128, 4
rng = np.random.default_rng(0)
rowptr, col = bench.make_ba_csr(N, 32, 0)
X = rng.normal(size=(N, d)).astype(np.float32)
sc = lambda *s_: (rng.normal(size=s_) * 0.3).astype(np.float32)
W = dict(W1=sc(d, 20), b1=sc(20), W2=sc(20, 20), b2=sc(20), W3=sc(20, 20), b3=sc(20), Wp=sc(Cc, 60), bp=sc(Cc))
eng = gnnx.Engine(0); eng.set_model(W)
eng.set_graph_csr(rowptr, col, X, rng.integers(0, Cc, N).astype(np.int32), rng.integers(0, Cc, N).astype(np.int32))
dbg = torch.zeros((1 << 19) + 64, dtype=torch.float32, device='cuda')
lib = _abi.lib(); lib.gx_debug_set_dump.argtypes = [C.c_void_p, C.c_void_p]
lib.gx_debug_set_dump(eng._h, C.c_void_p(dbg.data_ptr()))
names = ['F0', 'F1', 'F2', 'S', 'B2', 'B1', 'B0', 'P']
for k in sorted({1, K}):
    nodes = np.random.default_rng(1).permutation(N)[:k].astype(np.int32)
    eng.plan_nodes(nodes, 3, fetch=False)
    te = eng._plan_sizes[2]
    out = torch.empty(te, dtype=torch.float32, device='cuda')
    hp = eng.make_hparams(num_epochs=EP, init=_abi.GX_INIT_PHILOX, seed=1)
    eng.explain_nodes_ptr(hp, _abi.GX_DEVICE, 0, out.data_ptr()); torch.cuda.synchronize()
    D = dbg.cpu().numpy()[(1 << 19):(1 << 19) + 14]
    ph = D[:8] / max(EP - 1, 1)
    print('tasks %d  n,n1,n2,np_in,e_d,threads %s  kernel ms %.1f' % (k, D[8:14].astype(int), eng.last_explain_ms()))
    print('   kcycles/epoch ' + ' '.join('%s %.0f' % (nm, v / 1e3) for nm, v in zip(names, ph)) + '  total %.0f' % (ph.sum() / 1e3))
