"""Pure (no process group) checks of the rank-layout math; golden layouts from SURVEY.md 2.1 C02
(reference docstrings: dist/process_topo.py:72-90, Intro.md:33-45)."""
import pytest

from torchdistpackage_b200.dist.process_topo import (compute_axis_layout, compute_layout,
                                                     compute_moe_layout)
from torchdistpackage_b200.dist.node_group import node_rank_lists


def test_layout_data_pipe_tensor():
    lay = compute_layout(16, [("data", 4), ("pipe", 2), ("tensor", 2)])
    assert lay["tensor"] == [[i, i + 1] for i in range(0, 16, 2)]
    assert lay["pipe"] == [[0, 2], [4, 6], [8, 10], [12, 14], [1, 3], [5, 7], [9, 11], [13, 15]]
    assert lay["data"] == [[0, 4, 8, 12], [1, 5, 9, 13], [2, 6, 10, 14], [3, 7, 11, 15]]
    assert lay["model"] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]]


def test_layout_pipe_tensor_data():
    lay = compute_layout(16, [("pipe", 2), ("tensor", 2), ("data", 4)])
    assert lay["pipe"][0] == [0, 8]
    assert [0, 4] in lay["tensor"] and [8, 12] in lay["tensor"]
    assert lay["data"] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]]
    assert lay["model"][0] == [0, 4, 8, 12]


def test_layout_bench_config5():
    lay = compute_layout(8, [("data", 2), ("pipe", 2), ("tensor", 2)])
    assert lay["tensor"] == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert lay["pipe"] == [[0, 2], [4, 6], [1, 3], [5, 7]]
    assert lay["data"] == [[0, 4], [1, 5], [2, 6], [3, 7]]


def test_every_rank_in_exactly_one_group_per_axis():
    for cfg in ([("data", 2), ("pipe", 3), ("tensor", 4)], [("tensor", 8)], [("data", 8)]):
        world = 1
        for _, s in cfg:
            world *= s
        lay = compute_layout(world, cfg)
        for name, _ in cfg:
            flat = sorted(r for g in lay[name] for r in g)
            assert flat == list(range(world))


def test_layout_rejects_bad_product():
    with pytest.raises(ValueError):
        compute_layout(8, [("data", 3), ("tensor", 2)])


def test_axis_layout_stride():
    assert compute_axis_layout(8, 2, [2]) == [[0, 2], [4, 6], [1, 3], [5, 7]]
    assert compute_axis_layout(8, 4, []) == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_moe_layout():
    ep, dp, e, d = compute_moe_layout([list(range(8))], moe_ep_size=4)
    assert (e, d) == (4, 2)
    assert ep == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert dp == [[0, 4], [1, 5], [2, 6], [3, 7]]
    ep2, dp2, _, _ = compute_moe_layout([[0, 2, 4, 6], [1, 3, 5, 7]], moe_dp_size=2)
    assert ep2 == [[0, 2], [4, 6], [1, 3], [5, 7]]
    assert dp2 == [[0, 4], [2, 6], [1, 5], [3, 7]]
    with pytest.raises(ValueError):
        compute_moe_layout([list(range(8))], moe_dp_size=3, moe_ep_size=2)


def test_node_rank_lists():
    assert node_rank_lists(8, 8) is None            # single node: no hybrid split
    assert node_rank_lists(12, 8) is None
    assert node_rank_lists(16, 8) == [list(range(8)), list(range(8, 16))]
