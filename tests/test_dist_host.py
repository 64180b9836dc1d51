"""CPU: host-side bookkeeping of gnnx.dist (no GPU, no process group): shard layout invariants and the per-graph k-hop size cache."""
import types

import numpy as np

import conftest  # noqa: F401  (sys.path)
from gnnx import dist as gdist


def test_shard_layout_invariants():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        sizes = rng.integers(1, 5000, 137)
        shards, slot, src_off, offsets = gdist.shard_layout(sizes, world)
        allpos = np.sort(np.concatenate(shards))
        assert np.array_equal(allpos, np.arange(len(sizes)))                       # every item owned exactly once
        assert all(np.all(np.diff(s) > 0) for s in shards if len(s) > 1)           # each shard ascending (= the order a rank packs its items in)
        payload = [int(sizes[s].sum()) for s in shards]
        assert slot == max(payload)
        assert offsets[0] == 0 and np.array_equal(np.diff(offsets), sizes)
        for r, s in enumerate(shards):                                             # items of a rank sit back to back inside the rank's slot
            o = r * slot
            for p in s:
                assert src_off[p] == o
                o += sizes[p]
        # cost balance: dealing the descending order round-robin keeps the ranks within one largest item of each other
        assert max(payload) - min(payload) <= sizes.max()
        # identical on every rank = a pure function of its inputs
        again = gdist.shard_layout(sizes.copy(), world)
        assert all(np.array_equal(a, b) for a, b in zip(shards, again[0])) and slot == again[1]


def test_shard_layout_fewer_items_than_ranks():
    shards, slot, src_off, offsets = gdist.shard_layout(np.array([5, 3]), 4)
    assert sorted(len(s) for s in shards) == [0, 0, 1, 1] and slot == 5 and offsets[-1] == 8


class _FakeEngine:
    def __init__(self, n):
        self.num_nodes = n
        self.calls = []

    def count_nodes(self, nodes, hops):
        self.calls.append(np.array(nodes))
        nodes = np.asarray(nodes, np.int64)
        return nodes * 2 + hops, nodes * 10


def test_count_nodes_cached_asks_the_device_once_per_node():
    eng = _FakeEngine(50)
    ex = types.SimpleNamespace(engine=eng, n_hops=3, _current_graph=0)
    nodes = np.array([4, 9, 4, 30], np.int64)
    n, e = gdist.count_nodes_cached(ex, nodes)
    assert np.array_equal(n, nodes * 2 + 3) and np.array_equal(e, nodes * 10)
    assert len(eng.calls) == 1 and np.array_equal(eng.calls[0], [4, 9, 30])        # duplicates folded, one call
    n2, e2 = gdist.count_nodes_cached(ex, np.array([9, 30, 9]))
    assert len(eng.calls) == 1 and np.array_equal(n2, [21, 63, 21])                # served from the cache
    gdist.count_nodes_cached(ex, np.array([9, 31]))
    assert len(eng.calls) == 2 and np.array_equal(eng.calls[1], [31])              # only the node not seen yet
    ex.n_hops = 2                                                                   # another hop count = another cache line
    n3, _ = gdist.count_nodes_cached(ex, np.array([9]))
    assert len(eng.calls) == 3 and n3[0] == 9 * 2 + 2
    ex._current_graph = 1                                                           # another graph of a multi-graph Explainer
    gdist.count_nodes_cached(ex, np.array([9]))
    assert len(eng.calls) == 4
