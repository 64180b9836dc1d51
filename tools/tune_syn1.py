"""Sweep the scheduling knobs of the 700-node syn1 batch (one process per setting: the knobs are read once):
GNNX_EXCLUSIVE_TOPK (tasks of the 2-per-SM class that get an SM of their own) and GNNX_CLASS_THREADS (threads per launch class)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, os.path.join(%r, "gnn-model-explainer_b200")); sys.path.insert(0, %r)
import bench, gnnx
from gnnx import _abi
g = bench.load_syn1("syn1")
eng = gnnx.Engine(0); eng.set_model(g["weights"]); eng.set_graph_csr(g["rowptr"], g["col"], g["feat"], g["label"], g["pred_label"])
eng.plan_nodes(np.arange(g["N"], dtype=np.int32), 3, fetch=False)
out = torch.empty(eng._plan_sizes[2], dtype=torch.float32, device="cuda")
hp = eng.make_hparams(init=_abi.GX_INIT_PHILOX, seed=3)
ms = []
for _ in range(12):
    eng.explain_nodes_ptr(hp, _abi.GX_DEVICE, 0, out.data_ptr()); torch.cuda.synchronize(); ms.append(eng.last_explain_ms())
print(json.dumps({"min": min(ms), "med": float(np.median(ms))}))
''' % (ROOT, ROOT)

rows = []
for topk in ("0", "6", "12", "20", "30", "38"):
    for threads in ("", "128,256,256,512,512", "128,128,256,512,512", "128,256,256,384,512", "64,128,256,512,512", "128,256,512,512,512"):
        env = dict(os.environ, GNNX_EXCLUSIVE_TOPK=topk)
        if threads:
            env["GNNX_CLASS_THREADS"] = threads
        elif topk not in ("12",):
            pass
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = {"error": (r.stderr or r.stdout)[-200:]}
        d.update(topk=int(topk), class_threads=threads or "default")
        rows.append(d)
        print(json.dumps(d), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_syn1.json"), "w"), indent=1)
