"""Per-task timeline of one explain batch (debug hook): start/end (globaltimer ns), SM id, threads."""
import sys, os, numpy as np, ctypes as C, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import conftest, util  # noqa
from gnnx import _abi
fx = util.load_fixture('syn1')
eng = util.make_engine(fx)
dbg = torch.zeros((1 << 19) + 64 + 6 * 800, dtype=torch.float32, device='cuda')
lib = _abi.lib(); lib.gx_debug_set_dump.argtypes = [C.c_void_p, C.c_void_p]
nodes = list(range(700))
pl = eng.plan_nodes(nodes, 3); o = np.zeros(pl.total_edges, np.float32)
hp = eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=1)
eng.explain_nodes_host(hp, None, o)   # warm
lib.gx_debug_set_dump(eng._h, C.c_void_p(dbg.data_ptr()))
eng.explain_nodes_host(hp, None, o)
raw = dbg.cpu().numpy()[(1 << 19) + 64:].view(np.uint64)[:3 * 700].reshape(700, 3)
st, en = raw[:, 0].astype(np.int64), raw[:, 1].astype(np.int64)
smid = (raw[:, 2] >> np.uint64(32)).astype(int); thr = (raw[:, 2] & np.uint64(0xffffffff)).astype(int)
t0 = st.min(); st = (st - t0) / 1e6; en = (en - t0) / 1e6
E = np.diff(pl.edge_off); N = np.diff(pl.node_off)
print('makespan %.3f ms; kernel ms %.3f' % (en.max(), eng.last_explain_ms()))
order = np.argsort(-en)
print('last finishers: task n E thr sm start end dur')
for t in order[:12]:
    print('  ', t, N[t], E[t], thr[t], smid[t], '%.3f %.3f %.3f' % (st[t], en[t], en[t] - st[t]))
for th in sorted(set(thr)):
    m = thr == th
    print('threads', th, 'tasks', m.sum(), 'dur mean %.3f max %.3f; start max %.3f; end max %.3f' % ((en - st)[m].mean(), (en - st)[m].max(), st[m].max(), en[m].max()))
# per-SM busy
for s_ in list(np.argsort([-(en[smid == k].max() if (smid == k).any() else 0) for k in range(148)])[:5]):
    m = smid == s_
    print('SM', s_, 'tasks', m.sum(), 'last end %.3f' % en[m].max(), 'durs', np.round(np.sort((en - st)[m])[::-1][:6], 2))
