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


# ------------------------------------------------------------------------------------------------
# denoise_graph: the step every consumer of the masks runs next (explain.py:238-288,308; utils/io_utils.py:193-245)
# ------------------------------------------------------------------------------------------------
def graph_from_edges(num_nodes, node_idx, ei, ej, w, feat=None, label=None, max_component=True):
    """networkx graph of the thresholded edges, exactly the object denoise_graph returns (io_utils.py:207-245): all
    `num_nodes` nodes exist first (node_idx carries self=1, optional feat/label attributes), then either the largest
    connected component or the non-isolated nodes are kept."""
    import networkx as nx
    G = nx.Graph()
    G.add_nodes_from(range(num_nodes))
    G.nodes[node_idx]["self"] = 1
    if feat is not None:
        for node in G.nodes():
            G.nodes[node]["feat"] = feat[node]
    if label is not None:
        for node in G.nodes():
            G.nodes[node]["label"] = label[node]
    G.add_weighted_edges_from(zip((int(i) for i in ei), (int(j) for j in ej), (float(x) for x in w)))
    if max_component:
        largest_cc = max(nx.connected_components(G), key=len)
        return G.subgraph(largest_cc).copy()
    G.remove_nodes_from(list(nx.isolates(G)))
    return G


def denoise_graph(adj, node_idx, feat=None, label=None, threshold=None, threshold_num=None, max_component=True):
    """Same signature and result as the reference's io_utils.denoise_graph for a DENSE (n,n) mask (API compatibility for
    callers that hold dense arrays).  Explainer.denoise_nodes is the batched path: the threshold select and the edge
    compaction run on device on the packed masks (gx_denoise_topk)."""
    adj = np.asarray(adj)
    n = adj.shape[-1]
    if threshold_num is not None:
        pos = adj[adj > 0]
        k = min(len(pos), 2 * threshold_num)          # symmetric: every edge appears twice
        threshold = np.partition(pos, len(pos) - k)[len(pos) - k]
    ei, ej = np.nonzero(adj >= threshold) if threshold is not None else np.nonzero(adj > 1e-6)
    return graph_from_edges(n, node_idx, ei, ej, adj[ei, ej], feat, label, max_component)
