import sys, os, numpy as np, ctypes as C, torch
sys.path.insert(0,'tests'); import conftest, util
import gnnx_oracle as O
from gnnx import _abi
name=os.environ.get('FIX','syn4')
fx = util.load_fixture(name)
eng = util.make_engine(fx)
node=int(os.environ.get('NODE',str(fx.nodes[-1])))
d=fx.feat.shape[1]
idx,srp,scol,sfeat,slabel,nbrs = O.extract_neighborhood(fx.rowptr,fx.col,fx.feat,fx.label,node,3)
A=O.dense_from_csr(srp,scol); n=len(nbrs); ei,ej=np.nonzero(A)
M0=np.zeros((n,n),np.float32); M0[ei,ej]=fx.gold['n%d_m0'%node]
deg=A.sum(1).astype(int)
from collections import deque
dist=-np.ones(n,int); dist[idx]=0; dq=deque([idx])
while dq:
    u=dq.popleft()
    for v in scol[srp[u]:srp[u+1]]:
        if dist[v]<0: dist[v]=dist[u]+1; dq.append(v)
order=sorted(range(n), key=lambda c:(dist[c],-deg[c],c))   # level order (canonical ids)
order=np.array(order); n1=(dist<=1).sum(); n2=(dist<=2).sum()
W=fx.weights; f=np.float64
S=1/(1+np.exp(-M0.astype(f))); a=A*(S+S.T)/2
X=sfeat.astype(f); sF=np.full(d,0.5)
H=[X*sF]; Yh=[]; q=[]
Ws=[W['W1'].astype(f),W['W2'].astype(f),W['W3'].astype(f)]; bs=[W['b1'].astype(f),W['b2'].astype(f),W['b3'].astype(f)]
for l in range(3):
    Y=(a@H[-1])@Ws[l]+bs[l]; ql=np.maximum(np.sqrt((Y*Y).sum(1,keepdims=True)),1e-12); Yl=Y/ql
    Yh.append(Yl); q.append(ql); H.append(np.maximum(Yl,0) if l<2 else Yl)
emb=np.concatenate([H[1][idx],H[2][idx],H[3][idx]]); logits=W['Wp'].astype(f)@emb+W['bp']; p=np.exp(logits-logits.max()); p/=p.sum()
g=p.copy(); g[int(slabel[idx])]-=1; dEmb=W['Wp'].astype(f).T@g
dE=[np.zeros((n,20)) for _ in range(3)]
for l in range(3): dE[l][idx]=dEmb[20*l:20*l+20]
dH=np.zeros((n,20)); dZs=[None]*3
for l in (2,1,0):
    dYh=dE[l]+dH
    if l<2: dYh=dYh*(Yh[l]>0)
    dY=(dYh-Yh[l]*(Yh[l]*dYh).sum(1,keepdims=True))/q[l]; dZ=dY@Ws[l].T; dZs[l]=dZ; dH=a.T@dZ
U=a@X
dbg=torch.zeros(1<<20,dtype=torch.float32,device='cuda')
lib=_abi.lib(); lib.gx_debug_set_dump.argtypes=[C.c_void_p,C.c_void_p]
lib.gx_debug_set_dump(eng._h, C.c_void_p(dbg.data_ptr()))
pl = eng.plan_nodes([node],3); o=np.zeros(pl.total_edges,np.float32)
eng.explain_nodes_host(eng.make_hparams(num_epochs=3), util.golden_m0(fx,pl), o)
D=dbg.cpu().numpy(); hd=D[:16].astype(int); slab=D[16:16+hd[0]]
dp=hd[14]; nw=hd[15]
def arr(off,rows,cols,stride=None):
    stride=stride or cols
    return slab[off:off+rows*stride].reshape(rows,stride)[:,:cols]
def cmp(nm,got,ref):
    err=np.abs(got-ref).max(1); print('%-6s max err %.2e'%(nm,err.max()),' rows(lo) with err>1e-5:',np.nonzero(err>1e-5)[0][:12], 'levels', dist[order][np.nonzero(err>1e-5)[0][:12]])
print('n',n,'n1',n1,'n2',n2,'deg(lo order)',deg[order][:12])
cmp('U',arr(hd[2],n2,d,dp),U[order[:n2]])
cmp('Yh1',arr(hd[3],n2,20),Yh[0][order[:n2]])
cmp('Yh2',arr(hd[5],n1,20),Yh[1][order[:n1]])
cmp('dZ2',arr(hd[7],n1,20),dZs[1][order[:n1]])
cmp('dZ1s',arr(hd[8],n2,d,dp),(dZs[0]*sF)[order[:n2]])
cmp('dZ3',arr(hd[11],1,20),dZs[2][order[:1]])
gF=arr(hd[9],nw,d,dp).sum(0); print('gF got',gF[:6],'ref',(dZs[0]*U).sum(0)[:6])
