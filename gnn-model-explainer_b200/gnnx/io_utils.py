"""TU-Dortmund graph-kernel dataset reader for graph-classification mode (the format of Mutagenicity, BASELINE
config 4).  Host I/O only; mirrors the semantics of the reference's utils/io_utils.py:read_graphfile (426-562)
followed by graph_utils.GraphSampler's padding (graph_utils.py:31-60,128-145) for feature_type 'node-label':
  * graph labels are renumbered in order of first appearance in <name>_graph_labels.txt;
  * node labels are shifted by their minimum and one-hot encoded (the node features);
  * a graph's nodes are those that occur in its edge lines, numbered in order of first appearance
    (nx.from_edgelist insertion order); nodes without edges do not exist;
  * graphs with more than max_nodes nodes are dropped; the others are zero-padded to max_nodes.
Returns the arrays Explainer(..., graph_mode=True) takes."""
import os

import numpy as np


def read_tu_dataset(datadir, dataname, max_nodes=100):
    prefix = os.path.join(datadir, dataname, dataname)
    graph_of_node = np.loadtxt(prefix + "_graph_indicator.txt", dtype=np.int64, ndmin=1)          # 1-based graph ids
    node_labels = None
    if os.path.exists(prefix + "_node_labels.txt"):
        node_labels = np.loadtxt(prefix + "_node_labels.txt", dtype=np.int64, ndmin=1)
        node_labels = node_labels - node_labels.min()
        d = int(node_labels.max()) + 1
    raw_labels = np.loadtxt(prefix + "_graph_labels.txt", dtype=np.int64, ndmin=1)
    order = {}
    for v in raw_labels:
        order.setdefault(int(v), len(order))
    graph_labels = np.array([order[int(v)] for v in raw_labels], dtype=np.int64)
    G_all = len(raw_labels)
    edges = np.loadtxt(prefix + "_A.txt", dtype=np.int64, delimiter=",", ndmin=2)                   # 1-based node ids
    per_graph = [[] for _ in range(G_all)]
    for e0, e1 in edges:
        per_graph[graph_of_node[e0 - 1] - 1].append((int(e0), int(e1)))
    adjs, feats, labels, sizes = [], [], [], []
    for g in range(G_all):
        ids = {}
        for e0, e1 in per_graph[g]:
            ids.setdefault(e0, len(ids)); ids.setdefault(e1, len(ids))
        n = len(ids)
        if n == 0 or (max_nodes is not None and n > max_nodes):
            continue
        A = np.zeros((max_nodes, max_nodes), dtype=np.uint8)
        for e0, e1 in per_graph[g]:
            if e0 != e1:
                A[ids[e0], ids[e1]] = 1; A[ids[e1], ids[e0]] = 1
        adjs.append(A)
        if node_labels is not None:
            X = np.zeros((max_nodes, d), dtype=np.float32)
            for u, k in ids.items():
                X[k, node_labels[u - 1]] = 1.0
            feats.append(X)
        labels.append(graph_labels[g]); sizes.append(n)
    out = dict(adj=np.stack(adjs), label=np.asarray(labels, np.int64), num_nodes=np.asarray(sizes, np.int64))
    if node_labels is not None:
        out["feat"] = np.stack(feats)
    return out
