"""CPU, world_size 2, gloo: the N>1 path (deal nodes across ranks, ONE ragged all-gather of the packed
masks, reassembly in input order) must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _payload(item):
    rng = np.random.default_rng(1000 + item)
    return rng.random(3 + (item * 7) % 11).astype(np.float32)


def _worker(rank, world, port, num_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest  # noqa: F401  (sys.path)
    from gnnx.dist import shard_layout, allgather_packed
    sizes_all = np.array([len(_payload(i)) for i in range(num_items)], np.int64)      # known on every rank (gx_count_nodes on the GPU path)
    pos = shard_layout(sizes_all, world)[0][rank]
    vals = np.concatenate([_payload(i) for i in pos]) if len(pos) else np.zeros(0, np.float32)
    calls = []
    orig = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    values, offsets = allgather_packed(torch.from_numpy(vals), sizes_all, rank, world)
    assert len(calls) == 1, "the exchange must be ONE collective"
    q.put((rank, values.numpy(), np.asarray(offsets)))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, num_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_items, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    expect = np.concatenate([_payload(i) for i in range(num_items)])
    exp_off = np.concatenate([[0], np.cumsum([len(_payload(i)) for i in range(num_items)])])
    for rank, values, offsets in res:
        assert np.array_equal(offsets, exp_off), "rank %d offsets" % rank
        assert np.array_equal(values, expect), "rank %d values differ from the single-process order" % rank


def test_allgather_packed_world2():
    _run(2, 23)


def test_allgather_packed_world2_fewer_items_than_ranks():
    _run(2, 1)
