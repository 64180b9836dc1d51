#!/usr/bin/env python
"""Host-side cost of gnnx.dist.explain_nodes_sharded per step, measured on ONE GPU with a world-size-1 process group: the list handling
(k-hop size cache, shard layout memo) is the same work every rank of an N-rank run does for the whole list, so its cost here is the
per-step host overhead of the N-GPU weak-scaling line.    python tools/dist_overhead.py [list multiple, default 8]"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-model-explainer_b200"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from gnnx.dist import explain_nodes_sharded, ensure_comm, count_nodes_cached  # noqa: E402


def main():
    mult = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
    g = bench.load_syn1("syn1")
    ex = bench.make_explainer(g, 0, init="device")
    ensure_comm(ex.engine)
    rows = []
    for m in (1, mult):
        nodes = np.tile(np.arange(g["N"], dtype=np.int32), m)
        for _ in range(3):
            explain_nodes_sharded(ex, nodes)
        torch.cuda.synchronize()
        wall, kern = [], []
        for _ in range(10):
            t0 = time.perf_counter()
            explain_nodes_sharded(ex, nodes)
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            kern.append(ex.engine.last_explain_ms())
        # the list handling alone (what every rank repeats for the WHOLE list)
        t0 = time.perf_counter()
        for _ in range(50):
            count_nodes_cached(ex, nodes)
        t_count = (time.perf_counter() - t0) / 50 * 1e3
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(10):
            explain_nodes_sharded(ex, nodes)
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
        row = {"nodes": int(len(nodes)), "wall_ms_med": float(np.median(wall)), "kernel_ms_med": float(np.median(kern)),
               "outside_kernel_ms": float(np.median(wall) - np.median(kern)), "count_nodes_cached_ms": t_count}
        print(json.dumps(row), flush=True)
        print(s.getvalue()[:3000], flush=True)
        rows.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "dist_overhead.json"), "w"), indent=1)
    ex.engine.comm_destroy(); ex.engine.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
