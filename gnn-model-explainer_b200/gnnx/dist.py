"""Multi-GPU plumbing: nodes-to-explain are independent units, so they are dealt across ranks
(one process per GPU) with NO data-path collective; the only exchange is ONE all-gather of the
packed edge masks at the end (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).

Per-node arithmetic never crosses a GPU, so results are bit-identical to the 1-GPU run
(tests/test_gpu_parity.py::test_sharding_is_bit_identical, tests/test_dist_gloo.py)."""
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


def allgather_packed(local_vals, local_sizes, positions, num_items, device=None, group=None):
    """All-gather ragged per-item float32 payloads.

    local_vals : 1-D float32 tensor, concatenation of this rank's item payloads
    local_sizes: 1-D int64 tensor, payload length per local item
    positions  : 1-D int64 tensor, global position of each local item
    Returns (values, offsets): values = payloads of ALL items concatenated in global position
    order (1-D float32 tensor on `device`), offsets int64[num_items+1]."""
    world = dist.get_world_size(group)
    device = device if device is not None else local_vals.device
    local_vals = local_vals.to(device=device, dtype=torch.float32).contiguous()
    local_sizes = local_sizes.to(device=device, dtype=torch.int64)
    positions = positions.to(device=device, dtype=torch.int64)
    # 1) tiny metadata exchange: how many items / payload floats each rank holds
    meta = torch.tensor([local_sizes.numel(), local_vals.numel()], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu()
    max_items, max_vals = int(metas[:, 0].max()), int(metas[:, 1].max())
    # 2) item tables (position, size), padded to the per-rank maximum
    tab = torch.full((max_items, 2), -1, dtype=torch.int64, device=device)
    tab[: local_sizes.numel(), 0] = positions
    tab[: local_sizes.numel(), 1] = local_sizes
    tabs = torch.empty((world, max_items, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(tabs.view(-1), tab.view(-1), group=group)
    # 3) THE all-gather of the masks (padded to the per-rank maximum payload)
    pay = torch.zeros(max_vals, dtype=torch.float32, device=device)
    pay[: local_vals.numel()] = local_vals
    pays = torch.empty((world, max_vals), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(pays.view(-1), pay, group=group)
    # 4) reorder into global position order (pure indexing)
    sizes = torch.zeros(num_items, dtype=torch.int64, device=device)
    src_rank = torch.zeros(num_items, dtype=torch.int64, device=device)
    src_off = torch.zeros(num_items, dtype=torch.int64, device=device)
    for r in range(world):
        k = int(metas[r, 0])
        if k == 0:
            continue
        pos_r, sz_r = tabs[r, :k, 0], tabs[r, :k, 1]
        sizes[pos_r] = sz_r
        src_rank[pos_r] = r
        src_off[pos_r] = torch.cumsum(sz_r, 0) - sz_r
    offsets = torch.zeros(num_items + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(sizes, 0)
    total = int(offsets[-1])
    item_of = torch.repeat_interleave(torch.arange(num_items, device=device), sizes, output_size=total)
    within = torch.arange(total, device=device) - offsets[:-1][item_of]
    values = pays[src_rank[item_of], src_off[item_of] + within]
    return values, offsets


def explain_nodes_sharded(explainer, node_indices, costs=None, group=None):
    """Explainer.explain_nodes across all ranks of the default process group.
    Every rank returns the packed masks of ALL nodes: (values float32 tensor, offsets int64
    tensor, local (plan, positions)); values[offsets[t]:offsets[t+1]] are the masked_adj entries
    of node_indices[t] at the row-major sub-adjacency slots."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nodes = np.asarray(node_indices)
    pos = shard_indices(len(nodes), world, rank, costs)
    dev = torch.device("cuda", explainer.engine.device) if torch.cuda.is_available() else torch.device("cpu")
    if len(pos):
        plan, edge_mask = explainer.explain_nodes_packed(nodes[pos])
        sizes = torch.from_numpy(np.diff(plan.edge_off).astype(np.int64))
        vals = torch.from_numpy(edge_mask)
    else:
        plan, sizes, vals = None, torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.float32)
    values, offsets = allgather_packed(vals, sizes, torch.from_numpy(pos.astype(np.int64)), len(nodes),
                                       device=dev, group=group)
    return values, offsets, (plan, pos)
