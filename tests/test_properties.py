"""Property tests (hypothesis) of the pure planning code: rank layouts, pipeline partitions,
parameter sharding, MoE slot plans, bucket layout.  These are the functions every rank must
evaluate identically without communicating, so their invariants are worth more than examples."""
import math

import torch
from hypothesis import given, settings, strategies as st

from torchdistpackage_b200.dist.process_topo import compute_layout, compute_moe_layout
from torchdistpackage_b200.dist.node_group import node_rank_lists, inter_node_rank_lists
from torchdistpackage_b200.parallel.pipeline_parallel.pipeline_helper import uniform_bounds, balanced_bounds
from torchdistpackage_b200.utils import greedy_partition_sizes

AXES = ["data", "pipe", "tensor"]


@settings(max_examples=120, deadline=None)
@given(st.permutations(AXES), st.lists(st.sampled_from([1, 2, 3, 4]), min_size=3, max_size=3))
def test_layout_invariants(order, sizes):
    config = list(zip(order, sizes))
    world = math.prod(sizes)
    lay = compute_layout(world, config)
    for axis, size in config:
        groups = lay[axis]
        assert all(len(g) == size for g in groups)
        assert sorted(r for g in groups for r in g) == list(range(world))       # a partition
        stride = math.prod(s for a, s in config[[a for a, _ in config].index(axis) + 1:])
        assert all(g[i + 1] - g[i] == stride for g in groups for i in range(size - 1))
    # two ranks share a model group iff they are at the same position of their data groups
    pos = {r: g.index(r) for g in lay["data"] for r in g}
    for g in lay["model"]:
        assert len({pos[r] for r in g}) == 1
    assert sorted(r for g in lay["model"] for r in g) == list(range(world))
    # the innermost axis is made of consecutive ranks
    inner = config[-1][0]
    assert all(g == list(range(g[0], g[0] + len(g))) for g in lay[inner])


@settings(max_examples=80, deadline=None)
@given(st.sampled_from([2, 4, 6, 8, 12, 16]), st.data())
def test_moe_split_invariants(dp, data):
    ep = data.draw(st.sampled_from([d for d in range(1, dp + 1) if dp % d == 0]))
    groups = [list(range(b, b + dp)) for b in (0, dp)]                         # two data groups
    ep_groups, dp_groups, ep_, mdp = compute_moe_layout(groups, moe_ep_size=ep)
    assert ep_ == ep and ep * mdp == dp
    for members in (ep_groups, dp_groups):
        assert sorted(r for g in members for r in g) == list(range(2 * dp))
    assert all(len(g) == ep for g in ep_groups) and all(len(g) == mdp for g in dp_groups)
    # an expert-parallel group and a replica group intersect in exactly one rank of a data group
    for e in ep_groups:
        for d in dp_groups:
            if (e[0] < dp) == (d[0] < dp):
                assert len(set(e) & set(d)) == 1


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 40), st.integers(1, 8), st.integers(0, 3))
def test_uniform_bounds_cover(n_items, parts, extra):
    b = uniform_bounds(n_items, parts, extra)
    assert len(b) == parts and b[0][0] == 0 and b[-1][1] == n_items
    assert all(lo <= hi for lo, hi in b)
    assert all(b[i][1] == b[i + 1][0] or b[i + 1][0] >= n_items for i in range(parts - 1))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 1000), min_size=1, max_size=24), st.integers(1, 6))
def test_balanced_bounds_are_optimal_contiguous_partitions(weights, parts):
    parts = min(parts, len(weights))
    b = balanced_bounds(weights, parts)
    assert len(b) == parts and b[0][0] == 0 and b[-1][1] == len(weights)
    assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
    assert all(hi > lo for lo, hi in b)                                         # no empty stage
    load = max(sum(weights[lo:hi]) for lo, hi in b)

    # brute force: the best achievable maximum stage load of a contiguous split into `parts`
    from functools import lru_cache
    pre = [0]
    for w in weights:
        pre.append(pre[-1] + w)

    @lru_cache(None)
    def best(i, k):
        if k == 1:
            return pre[len(weights)] - pre[i]
        return min(max(pre[j] - pre[i], best(j, k - 1)) for j in range(i + 1, len(weights) - k + 2))
    assert load == best(0, parts)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(1, 500), min_size=1, max_size=30), st.integers(1, 8))
def test_greedy_partition_is_contiguous_and_complete(numels, parts):
    owners = greedy_partition_sizes(numels, parts)
    assert len(owners) == len(numels) and owners[0] == 0
    assert all(0 <= o < parts for o in owners)
    assert all(owners[i] <= owners[i + 1] <= owners[i] + 1 for i in range(len(owners) - 1))


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 8), st.integers(1, 8))
def test_node_rank_lists(nodes, per_node):
    world = nodes * per_node
    intra, inter = node_rank_lists(world, per_node), inter_node_rank_lists(world, per_node)
    if nodes == 1:
        assert intra is None and inter is None
        return
    assert sorted(r for g in intra for r in g) == list(range(world))
    assert sorted(r for g in inter for r in g) == list(range(world))
    assert all(len(set(r // per_node for r in g)) == 1 for g in intra)          # one node each
    assert all(len(set(r % per_node for r in g)) == 1 for g in inter)           # same local index
    assert all(len(set(a) & set(b)) == 1 for a in intra for b in inter)


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 64), st.integers(1, 3), st.sampled_from([2, 4, 8]), st.integers(1, 8), st.data())
def test_moe_slot_plan_is_a_partial_injection(tokens, k, n_exp, cap, data):
    """Every kept (token, k) pair gets its own slot on the rank that owns its expert; dropped
    pairs (over capacity) get -1; two source ranks can never collide (slot ranges are per source)."""
    from torchdistpackage_b200.moe.layer import _Plan
    ep = data.draw(st.sampled_from([d for d in (1, 2, 4, 8) if n_exp % d == 0]))
    k = min(k, n_exp)
    seed = data.draw(st.integers(0, 10 ** 6))
    plans = []
    for r in range(ep):
        g = torch.Generator().manual_seed(seed + r)
        idx = torch.stack([torch.randperm(n_exp, generator=g)[:k] for _ in range(tokens)])
        plans.append(_Plan(idx, n_exp, ep, r, cap))
    taken = set()
    for r, p in enumerate(plans):
        rows, dst = p.dst_row.tolist(), p.dst_rank.tolist()
        for (row, d, keep) in zip(rows, dst, p.keep.tolist()):
            assert (row >= 0) == keep
            if keep:
                assert 0 <= row < p.slots_per_rank and 0 <= d < ep
                assert (d, row) not in taken
                taken.add((d, row))
                local_e, rest = divmod(row, ep * cap)
                assert rest // cap == r                                         # my source lane
        per_expert = torch.bincount(p.dst_rank.long() * p.e_local +
                                    torch.div(p.dst_row.clamp_min(0), ep * cap, rounding_mode="floor"),
                                    weights=p.keep.float(), minlength=n_exp)
        assert float(per_expert.max()) <= cap


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 40), st.integers(1, 40), st.booleans()), min_size=1, max_size=10),
       st.sampled_from([2e-4, 1e-3, 5e-3, 25.0]), st.booleans())
def test_ddp_bucket_plan_invariants(layers, cap_mb, as_view):
    """Whatever the parameter shapes and the bucket size: every trainable parameter lives in exactly
    one bucket, slots are 512-byte aligned and disjoint, only an oversized parameter may exceed
    the cap (alone), buckets fill in reverse registration order, gradients written through
    ``p.grad`` land in the flat buffer."""
    import torch.nn as nn
    import torchdistpackage_b200 as tdp
    mods = []
    for fin, fout, bias in layers:
        mods.append(nn.Linear(fin, fout, bias=bias))
    model = nn.Sequential(*mods)
    ddp = tdp.NaiveDDP(model, gradient_as_bucket_view=as_view, bucket_cap_mb=cap_mb)
    try:
        red = ddp.reducer
        names = [n for n, p in model.named_parameters()]
        placed = [n for b in red.buckets for n in b.names]
        assert sorted(placed) == sorted(names)
        cap_bytes = int(cap_mb * 1024 * 1024)
        for b in red.buckets:
            spans = []
            for n in b.names:
                v = b.views[n]
                off = (v.data_ptr() - b.buffer.data_ptr())
                assert off % 512 == 0 and v.numel() == dict(model.named_parameters())[n].numel()
                spans.append((off, off + v.numel() * v.element_size()))
            spans.sort()
            assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
            assert spans[-1][1] <= b.buffer.numel() * b.buffer.element_size()
            if len(b.names) > 1:
                assert spans[-1][1] <= max(cap_bytes, 512 * len(b.names))
        # reverse registration order: the last layer's parameters are in the first bucket
        assert names[-1] in red.buckets[0].names
        x = torch.randn(3, layers[0][0])
        y = x
        for (fin, fout, _), m in zip(layers, mods):
            if y.shape[-1] != fin:
                y = torch.randn(3, fin)
            y = m(y)
        y.sum().backward()
        ddp.reduce_gradients()
        if as_view:
            for n, p in model.named_parameters():
                if p.grad is not None:
                    b = red.param_bucket[n]
                    assert p.grad.data_ptr() == b.views[n].data_ptr()
    finally:
        ddp.remove_hooks()
