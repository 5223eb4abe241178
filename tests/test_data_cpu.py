"""Host-side input pipeline: DistributedSampler-equivalent index sharding."""
import torch

import vds_b200  # noqa: F401
from vds_b200.data import batches, shard_indices


def test_shard_indices_partition_and_match_torch_sampler():
    from torch.utils.data.distributed import DistributedSampler
    n, world = 103, 4
    ds = list(range(n))
    for epoch in (0, 3):
        got = [shard_indices(n, r, world, seed=7, epoch=epoch) for r in range(world)]
        assert len({len(g) for g in got}) == 1 and len(got[0]) == (n + world - 1) // world
        assert set().union(*map(set, got)) == set(range(n))          # every sample seen (tail padded by wrap-around)
        for r in range(world):
            s = DistributedSampler(ds, num_replicas=world, rank=r, shuffle=True, seed=7)
            s.set_epoch(epoch)
            assert list(iter(s)) == got[r]
    # drop_last: disjoint, equal-sized, no padding
    got = [shard_indices(n, r, world, seed=1, drop_last=True) for r in range(world)]
    flat = sum(got, [])
    assert len(flat) == (n // world) * world and len(set(flat)) == len(flat)
    assert shard_indices(10, 1, 2, shuffle=False) == [1, 3, 5, 7, 9]
    assert shard_indices(0, 0, 2) == []


def test_batches():
    idx = list(range(10))
    assert batches(idx, 4) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert batches(idx, 4, drop_last=False)[-1] == [8, 9]


def test_prefetcher_refuses_cpu():
    import pytest
    from vds_b200.data import DevicePrefetcher
    with pytest.raises(RuntimeError):
        DevicePrefetcher([], device="cpu")


def test_attention_backward_query_split_model():
    """engine._attn_q_splits: pure host logic choosing how many ways the query range of an attention backward is split
    when there are fewer (kv tile, head, batch) items than SMs (cross-attention, Lk = 512)."""
    from vds_b200.engine import _attn_q_splits as qs
    assert qs(65, 2, 4, 65) == 1                      # self-attention at L = 8208: 520 items >= 148 SMs, never split
    s = qs(4, 2, 4, 65)                               # cross-attention of the debug model: 32 items
    assert 2 <= s <= 9 and (32 * s + 147) // 148 <= 2   # at most two waves (was 10 splits = 3 waves)
    for items in (1, 8, 32, 100, 147):
        for nq in (1, 3, 17, 65):
            v = qs(items, 1, 1, nq)
            assert 1 <= v <= max(1, 2 * nq)           # never more splits than 64-row sub-tiles
    assert qs(1, 1, 1, 1) <= 2


def test_attn_bwd_tail_plan_covers_every_pair_once_and_balances():
    """vds_attn_bwd_tail_plan (host-only C entry point): the pieces the CTA-pair attention backward cuts the pairs of its
    partly filled last wave into.  Every pair's query range must be covered exactly once, no piece may be tiny, the pieces
    come longest first, and the longest-first schedule on the SM pairs must beat both the unsplit launch and round 1's
    uniform split."""
    import ctypes
    import heapq

    import vds_b200  # noqa: F401
    from vds_b200 import lib
    L = lib.lib()
    buf = (ctypes.c_uint32 * 256)()
    for (P, nq, C) in [(42, 129, 74), (10, 129, 74), (1, 129, 74), (36, 129, 74), (20, 64, 74), (8, 33, 74), (73, 129, 74)]:
        n = L.vds_attn_bwd_tail_plan(P, nq, C, buf, 256)
        if n == 0:
            assert P == 73      # splitting a nearly full wave does not pay
            continue
        pieces = [(buf[i] & 1023, (buf[i] >> 10) & 2047, buf[i] >> 21) for i in range(n)]
        cover = {p: [] for p in range(P)}
        for pair, q0, cnt in pieces:
            assert 0 <= pair < P and cnt >= 8 and q0 + cnt <= nq
            cover[pair].append((q0, cnt))
        for p, segs in cover.items():
            segs.sort()
            pos = 0
            for q0, cnt in segs:
                assert q0 == pos
                pos += cnt
            assert pos == nq, (P, nq, p, segs)
            assert len(segs) <= 3 or nq / len(segs) >= 8
        assert [c for _, _, c in pieces] == sorted((c for _, _, c in pieces), reverse=True)
        # greedy hand-out in launch order (what the block scheduler does), 10 sub-tile units of overhead per piece
        slots = [0.0] * C
        heapq.heapify(slots)
        for _, _, cnt in pieces:
            heapq.heappush(slots, heapq.heappop(slots) + cnt + 10.0)
        makespan = max(slots)
        assert makespan < nq + 10.0                       # better than unsplit
        if (P, nq, C) == (42, 129, 74):                   # the debug-8k self-attention tail: uniform 3-way split = 2 x (43 + 10)
            assert makespan <= 95.0
    assert L.vds_attn_bwd_tail_plan(42, 129, 74, buf, 4) == 0      # too small a table: unsplit


def test_wgrad_split_model_reproduces_the_measured_optima():
    """engine.wgrad_splits: host-side cost model of the split-K factor of the wgrad GEMMs; the factors below are the
    measured optima of scripts/wgrad_splits_bench.py on a B200 (debug-8k, DiT-B and DiT-XL token counts)."""
    import vds_b200  # noqa: F401
    from vds_b200.engine import wgrad_splits as w
    assert [w(1536, 512, 16416), w(2048, 512, 16416), w(512, 2048, 16416), w(512, 512, 16416)] == [12, 9, 9, 9]
    assert [w(1536, 512, 2176), w(2048, 512, 2176), w(512, 2048, 2176), w(512, 512, 2176)] == [3, 2, 2, 9]
    assert [w(1536, 512, 4128), w(2048, 512, 4128), w(512, 512, 4128)] == [3, 2, 9]
    assert w(24576, 4096, 1024) == 1                  # the grouped context_kv wgrad: hundreds of tiles, nothing to split
    for args in [(128, 128, 64), (512, 512, 100), (3456, 1152, 4128)]:
        assert 1 <= w(*args) <= max(1, (args[2] + 63) // 64)
