"""Multi-GPU plumbing: nodes-to-explain are independent units, so they are dealt across ranks (one process per GPU) with
NO data-path collective; the only exchange is ONE all-gather of the packed edge masks at the end.

How one collective suffices: every rank knows the k-hop subgraph size of EVERY node of the list (gx_count_nodes: the
integer frontier expansion without building a plan, ~0.2 ms for 700 nodes, remembered per graph), so the shard assignment (cost balanced), every
rank's payload size and every item's offset are known everywhere without a metadata exchange.  Each rank pads its packed masks to
the largest per-rank payload, ONE all-gather moves the slots (NCCL over NVLink through the library's own communicator,
gx_allgather_masks; torch.distributed -- gloo in the CPU tests -- when no engine communicator exists), and a device kernel
(gx_unshard_masks) scatters the slots into input order.  The masks never leave the device between the explainer kernels and
the collective.

Per-node arithmetic never crosses a GPU, so results are bit-identical to the 1-GPU run (tests/test_gpu_dist.py,
tests/test_dist_gloo.py, bench.py "shard_bit_identical")."""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(num_items, world, rank, costs=None):
    """Positions (into the caller's node list) owned by `rank`.  With costs: sort by cost
    descending and deal round-robin (LPT-style balance); otherwise plain round-robin."""
    if costs is None:
        order = np.arange(num_items)
    else:
        order = np.argsort(-np.asarray(costs), kind="stable")
    return np.sort(order[rank::world])


def shard_layout(sizes_all, world, costs=None):
    """Everything a rank needs to know about the exchange, computed identically on every rank from the per-item sizes:
    shards[r] = positions of rank r; slot = floats per rank in the all-gather (largest payload); src_off[p] = where item p sits in
    the gathered [world*slot] buffer; offsets[p] = where it goes in input order."""
    sizes_all = np.asarray(sizes_all, np.int64)
    num = len(sizes_all)
    order = np.argsort(-np.asarray(sizes_all if costs is None else costs), kind="stable")     # one sort; shard r = every world-th item of it
    shards = [np.sort(order[r::world]) for r in range(world)]
    slot = max(1, max(int(sizes_all[s].sum()) for s in shards))
    offsets = np.concatenate([[0], np.cumsum(sizes_all)]).astype(np.int64)
    src_off = np.zeros(num, np.int64)
    for r, s in enumerate(shards):
        if len(s):
            src_off[s] = r * slot + np.concatenate([[0], np.cumsum(sizes_all[s])[:-1]])
    return shards, slot, src_off, offsets


def allgather_packed(local_vals, sizes_all, rank, world, costs=None, group=None, engine=None, out=None, layout=None):
    """ONE all-gather of ragged per-item float32 payloads.
    local_vals : 1-D float32 tensor = this rank's items (positions shard_layout(...)[0][rank], ascending) concatenated
    sizes_all  : payload length of EVERY item of the list (known on every rank: gx_count_nodes)
    Returns (values, offsets): all payloads concatenated in input order, int64 offsets[num_items+1]."""
    shards, slot, src_off, offsets = layout if layout is not None else shard_layout(sizes_all, world, costs)
    sizes_all = np.asarray(sizes_all, np.int64)
    device = local_vals.device
    total = int(offsets[-1])
    assert local_vals.numel() == int(sizes_all[shards[rank]].sum()), "local payload does not match the shard layout"
    if engine is not None and getattr(engine, "comm_world", None) == world and device.type == "cuda":
        gathered = engine.allgather_masks(local_vals.contiguous(), slot)                 # gx_allgather_masks: one ncclAllGather
        values = out if out is not None else torch.empty(max(total, 1), dtype=torch.float32, device=device)
        engine.unshard_masks(gathered, src_off, offsets[:-1], sizes_all, values)         # gx_unshard_masks: device scatter
        return values[:total], offsets
    pay = torch.zeros(slot, dtype=torch.float32, device=device)
    pay[: local_vals.numel()] = local_vals
    gathered = torch.empty(world * slot, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, pay, group=group)                               # THE collective (gloo / torch's NCCL)
    item_of = np.repeat(np.arange(len(sizes_all)), sizes_all)
    idx = torch.from_numpy(src_off[item_of] + (np.arange(total) - offsets[:-1][item_of])).to(device)
    return gathered[idx], offsets


def ensure_comm(engine, group=None):
    """Bootstraps the engine's own NCCL communicator (gx_comm_init) once per process group: rank 0 creates the id, one
    broadcast_object_list transports it."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if getattr(engine, "comm_world", None) == world:
        return
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(world, rank, box[0])


def count_nodes_cached(explainer, nodes):
    """(n, E_d) of every node of the list.  The k-hop sizes are a property of (graph, node, n_hops), so they are counted once per
    Explainer graph (gx_count_nodes on the nodes not seen yet) and remembered: a steady stream of explain calls pays nothing here."""
    eng = explainer.engine
    key = (getattr(explainer, "_current_graph", 0), int(explainer.n_hops))
    cache = explainer.__dict__.setdefault("_count_cache", {})
    if key not in cache:
        N = eng.num_nodes
        cache[key] = (np.full(N, -1, np.int64), np.zeros(N, np.int64))
    n_c, e_c = cache[key]
    nodes = np.asarray(nodes, np.int64)
    todo = np.unique(nodes[n_c[nodes] < 0])
    if len(todo):
        n_new, e_new = eng.count_nodes(todo.astype(np.int32), explainer.n_hops)
        n_c[todo] = n_new; e_c[todo] = e_new
    return n_c[nodes], e_c[nodes]


def explain_nodes_sharded(explainer, node_indices, costs=None, group=None, use_engine_comm=True):
    """Explainer.explain_nodes across all ranks of the default process group.
    Every rank returns the packed masks of ALL nodes: (values float32 tensor, offsets int64 array, (local plan or None, local positions));
    values[offsets[t]:offsets[t+1]] are the masked_adj entries of node_indices[t] at the row-major sub-adjacency slots."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    eng = explainer.engine
    nodes = np.asarray(node_indices)
    dev = torch.device("cuda", eng.device)
    n_all, e_all = count_nodes_cached(explainer, nodes)
    # the layout of a node list is a pure function of (list, world, costs): remembered for the list that was explained last
    key = (nodes.tobytes(), world, None if costs is None else np.asarray(costs).tobytes(), getattr(explainer, "_current_graph", 0))
    memo = explainer.__dict__.get("_layout_memo")
    if memo is None or memo[0] != key:
        memo = (key, shard_layout(e_all, world, costs))
        explainer._layout_memo = memo
    layout = memo[1]
    pos = layout[0][rank]
    hp, init = explainer._hparams()
    if len(pos):
        # the canonical sub-graph description comes back to the host only when the torch-compatible init needs it (M0 gather)
        plan = eng.plan_nodes(nodes[pos], explainer.n_hops, fetch=(init == "torch"))
        m0_dev = None
        if init == "torch":   # every rank walks the whole list so that torch's RNG is consumed exactly as one process would
            m0_dev = torch.from_numpy(explainer._draw_m0_subset(plan, n_all, pos)).to(dev)
        local = eng.explain_nodes_device(hp, m0_dev)
    else:
        plan, local = None, torch.zeros(0, dtype=torch.float32, device=dev)
    if use_engine_comm:
        ensure_comm(eng, group)
    values, offsets = allgather_packed(local, e_all, rank, world, costs, group=group, engine=eng if use_engine_comm else None, layout=layout)
    return values, offsets, (plan, pos)
