"""Per-phase cycle counters of the explainer kernel (debug hook gx_debug_set_dump): averages over the
epochs of the first task of the launch.  Usage on the GPU box: python tools/phase_timers.py"""
import sys, os, numpy as np, ctypes as C, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest, util  # noqa
from gnnx import _abi
fx = util.load_fixture('syn1')
eng = util.make_engine(fx)
dbg = torch.zeros((1 << 19) + 64, dtype=torch.float32, device='cuda')
lib = _abi.lib(); lib.gx_debug_set_dump.argtypes = [C.c_void_p, C.c_void_p]
lib.gx_debug_set_dump(eng._h, C.c_void_p(dbg.data_ptr()))
for nodes in ([450], [300], [3], [0], list(range(700))):
    pl = eng.plan_nodes(nodes, 3); o = np.zeros(pl.total_edges, np.float32)
    eng.explain_nodes_host(eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=1), None, o)
    D = dbg.cpu().numpy()[(1 << 19):(1 << 19) + 12]
    ph = D[:6] / 99.0
    print('nodes', nodes[:2], 'n,n1,n2,np_in,e1,thr', D[6:12].astype(int),
          'cycles/epoch F1 %.0f F2 %.0f S %.0f B2 %.0f B1 %.0f P %.0f total %.0f' % (*ph, ph.sum()), 'kernel ms %.3f' % eng.last_explain_ms())
