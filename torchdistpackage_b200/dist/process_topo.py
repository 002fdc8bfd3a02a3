"""Process topology: named parallel axes -> process groups, plus the query API every other
component uses (``tpc``).

Behavioural parity with the reference ``ProcessTopology`` (dist/process_topo.py:53-259):
``setup_process_groups([(axis, size), ...])`` lists axes outermost -> innermost, an automatic
``'model'`` group is built when ``tensor`` or ``pipe`` is present, ``build_moe_groups`` splits
every data group into contiguous ``moe_ep`` groups and strided ``moe_dp`` groups.

Design differences:
* the rank-layout math is a pure function (:func:`compute_axis_layout`, :func:`compute_layout`,
  :func:`compute_moe_layout`) so it is unit-testable without a process group and reusable by
  launch tooling;
* every group can lazily own a :class:`~torchdistpackage_b200.ops.symm.SymmGroup`
  (NVSwitch symmetric memory + multicast) via :meth:`ProcessTopology.get_symm_group`;
* ``reset()`` exists so tests can rebuild topologies.
"""
from __future__ import annotations

from collections import defaultdict
from datetime import timedelta
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

AxisConfig = Sequence[Tuple[str, int]]


# --------------------------------------------------------------------------------------------
# pure layout math
# --------------------------------------------------------------------------------------------
def compute_axis_layout(world_size: int, size: int, inner_sizes: Sequence[int]) -> List[List[int]]:
    """Rank lists of one axis of extent ``size`` whose inner (faster varying) axes have the
    extents ``inner_sizes``.  Members of a group are ``stride = prod(inner_sizes)`` apart.

    Group order matches the reference enumeration (process_topo.py:32-51): inner offset is the
    outer loop, outer block the inner loop -- e.g. world 16, pipe=2 with inner [2] gives
    ``[0,2],[4,6],[8,10],[12,14],[1,3],...``.
    """
    size = int(size)
    stride = 1
    for s in inner_sizes:
        stride *= int(s)
    if size <= 0 or world_size % (size * stride) != 0:
        raise ValueError(f"axis of size {size} with inner stride {stride} does not tile "
                         f"world_size={world_size}")
    span = size * stride
    groups = []
    for offset in range(stride):
        for base in range(0, world_size, span):
            groups.append([base + offset + j * stride for j in range(size)])
    return groups


def gen_inner_ranks(world_size: int, group_size: int) -> List[List[int]]:
    """Consecutive rank lists of an innermost axis (reference helper, process_topo.py:28-30)."""
    return compute_axis_layout(world_size, group_size, ())


def gen_groups(world_size: int, group_size: int, strides=None, hook=None) -> List[List[int]]:
    """Reference-style front end of :func:`compute_axis_layout` (process_topo.py:32-51): rank
    lists of an axis of extent ``group_size`` whose inner axes have the extents ``strides``;
    ``hook`` (if given) is called once per rank list, in enumeration order."""
    lists = compute_axis_layout(world_size, group_size, tuple(strides or ()))
    if hook is not None:
        for ranks in lists:
            hook(ranks)
    return lists


def compute_layout(world_size: int, config: AxisConfig) -> Dict[str, List[List[int]]]:
    """All rank lists for a ``[(axis, size), ...]`` config (outermost first) incl. ``'model'``."""
    names = [c[0] for c in config]
    sizes = [int(c[1]) for c in config]
    total = 1
    for s in sizes:
        total *= s
    if total != world_size:
        raise ValueError(f"product of axis sizes {sizes} = {total} != world_size {world_size}")
    if len(set(names)) != len(names):
        raise ValueError(f"duplicate axis names in {names}")
    layout: Dict[str, List[List[int]]] = {}
    for i, (name, size) in enumerate(zip(names, sizes)):
        layout[name] = compute_axis_layout(world_size, size, sizes[i + 1:])
    if ("tensor" in names or "pipe" in names) and "data" in names:
        data_groups = layout["data"]
        layout["model"] = [[g[i] for g in data_groups] for i in range(len(data_groups[0]))]
    elif "tensor" in names or "pipe" in names:
        layout["model"] = [list(range(world_size))]
    return layout


def compute_moe_layout(data_groups: List[List[int]], moe_dp_size: Optional[int] = None,
                       moe_ep_size: Optional[int] = None):
    """Split each data-parallel group into expert-parallel (contiguous members) and
    replicated-expert data-parallel (members strided by ``ep``) groups
    (reference: process_topo.py:118-143).  Returns ``(ep_groups, dp_groups, ep, dp)``."""
    dp_world = len(data_groups[0])
    if moe_dp_size and not moe_ep_size:
        moe_ep_size = dp_world // int(moe_dp_size)
    elif moe_ep_size and not moe_dp_size:
        moe_dp_size = dp_world // int(moe_ep_size)
    elif not moe_dp_size and not moe_ep_size:
        raise ValueError("build_moe_groups needs moe_dp_size and/or moe_ep_size")
    moe_dp_size, moe_ep_size = int(moe_dp_size), int(moe_ep_size)
    if moe_dp_size * moe_ep_size != dp_world:
        raise ValueError(f"moe_dp_size({moe_dp_size}) * moe_ep_size({moe_ep_size}) "
                         f"!= data-parallel size ({dp_world})")
    ep_groups, dp_groups = [], []
    for ranks in data_groups:
        for g in range(dp_world // moe_ep_size):
            ep_groups.append([ranks[i] for i in range(g * moe_ep_size, (g + 1) * moe_ep_size)])
        for g in range(dp_world // moe_dp_size):
            dp_groups.append([ranks[i] for i in range(g, dp_world, moe_ep_size)])
    return ep_groups, dp_groups, moe_ep_size, moe_dp_size


# --------------------------------------------------------------------------------------------
# the context object
# --------------------------------------------------------------------------------------------
class ProcessTopology:
    """Singleton-style registry ``{axis -> process group / my ranks / all rank lists}``."""

    _instance: Optional["ProcessTopology"] = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._init_state()
        return cls._instance

    def _init_state(self) -> None:
        self._groups: Dict[str, object] = {}
        self._ranks_in_group: Dict[str, List[int]] = {}
        self._ranks_all: Dict[str, List[List[int]]] = defaultdict(list)
        self._symm: Dict[str, object] = {}
        self._timeout = timedelta(seconds=100)
        self.verbose = True

    def reset(self) -> None:
        """Forget all groups (does not destroy the NCCL communicators)."""
        self._init_state()

    # ---------------------------------------------------------------- construction
    def _register(self, mode: str, all_rank_lists: List[List[int]]) -> None:
        rank = dist.get_rank()
        for ranks in all_rank_lists:
            self._ranks_all[mode].append(list(ranks))
            # new_group is collective over the whole world: every rank creates every group
            grp = dist.new_group(list(ranks), timeout=self._timeout)
            if rank in ranks:
                self._groups[mode] = grp
                self._ranks_in_group[mode] = list(ranks)
                if self.verbose and rank == ranks[0]:
                    print(f"[tpc] group {mode}: ranks {list(ranks)}", flush=True)

    def setup_process_groups(self, config: AxisConfig) -> None:
        """``config`` e.g. ``[('data', 4), ('pipe', 2), ('tensor', 2)]`` for 16 ranks."""
        world = dist.get_world_size()
        layout = compute_layout(world, config)
        self._groups["global"] = None
        self._ranks_in_group["global"] = list(range(world))
        self._ranks_all["global"] = [list(range(world))]
        for name, _ in config:
            self._register(name, layout[name])
        if "model" in layout:
            self._register("model", layout["model"])

    def build_moe_groups(self, moe_dp_size: Optional[int] = None,
                         moe_ep_size: Optional[int] = None) -> None:
        if "data" not in self._ranks_all:
            raise RuntimeError("build_moe_groups requires a 'data' axis")
        ep_groups, dp_groups, ep, dp = compute_moe_layout(self._ranks_all["data"], moe_dp_size,
                                                          moe_ep_size)
        if self.verbose and dist.get_rank() == 0:
            print(f"[tpc] MoE groups: moe_dp_size={dp}, moe_ep_size={ep}", flush=True)
        self._register("moe_ep", ep_groups)
        self._register("moe_dp", dp_groups)

    def setup_node_groups(self, num_per_node: int = 8):
        """Register intra-node groups as axis ``'node'`` (see dist/node_group.py)."""
        world = dist.get_world_size()
        if world % num_per_node != 0:
            return None
        lists = [list(range(n * num_per_node, (n + 1) * num_per_node))
                 for n in range(world // num_per_node)]
        self._register("node", lists)
        return self._groups.get("node")

    # ---------------------------------------------------------------- queries
    def _require(self, mode: str) -> None:
        if mode not in self._groups:
            raise AssertionError(f"{mode} is not initialized!")

    def get_group(self, mode: str):
        self._require(mode)
        return self._groups[mode]

    def get_group_rank(self, mode: str) -> int:
        self._require(mode)
        return self._ranks_in_group[mode].index(dist.get_rank())

    def get_ranks_in_group(self, mode: str) -> List[int]:
        self._require(mode)
        return self._ranks_in_group[mode]

    def get_group_size(self, mode: str) -> int:
        self._require(mode)
        return len(self._ranks_in_group[mode])

    def get_tp_rank(self): return self.get_group_rank("tensor")
    def get_pp_rank(self): return self.get_group_rank("pipe")
    def get_dp_rank(self): return self.get_group_rank("data")
    def get_mp_rank(self): return self.get_group_rank("model")
    def get_tp_size(self): return self.get_group_size("tensor")
    def get_pp_size(self): return self.get_group_size("pipe")
    def get_dp_size(self): return self.get_group_size("data")
    def get_mp_size(self): return self.get_group_size("model")

    def is_first_in_group(self, mode: str) -> bool:
        return self.get_group_rank(mode) == 0

    def is_last_in_group(self, mode: str) -> bool:
        return dist.get_rank() == self.get_ranks_in_group(mode)[-1]

    def is_first_in_tensor_group(self): return self.is_first_in_group("tensor")
    def is_last_in_tensor_group(self): return self.is_last_in_group("tensor")
    def is_first_in_pipeline_group(self): return self.is_first_in_group("pipe")
    def is_last_in_pipeline_group(self): return self.is_last_in_group("pipe")
    def is_first_in_data_group(self): return self.is_first_in_group("data")
    def is_last_in_data_group(self): return self.is_last_in_group("data")
    def is_first_in_model_group(self): return self.is_first_in_group("model")
    def is_last_in_model_group(self): return self.is_last_in_group("model")

    def get_prev_global_rank(self, mode: str = "pipe") -> int:
        ranks = self.get_ranks_in_group(mode)
        return ranks[(self.get_group_rank(mode) - 1) % len(ranks)]

    def get_next_global_rank(self, mode: str = "pipe") -> int:
        ranks = self.get_ranks_in_group(mode)
        return ranks[(self.get_group_rank(mode) + 1) % len(ranks)]

    def is_mode_inited(self, mode: str) -> bool:
        """Axis exists *and* actually spans more than one rank."""
        return mode in self._groups and self.get_group_size(mode) > 1

    def all_dp_ranks(self) -> List[List[int]]:
        return self._ranks_all["data"]

    def all_ranks(self, mode: str) -> List[List[int]]:
        self._require(mode)
        return self._ranks_all[mode]

    def is_first_group(self, mode: str) -> bool:
        self._require(mode)
        return self._ranks_in_group[mode] == self._ranks_all[mode][0]

    # ---------------------------------------------------------------- symmetric memory
    def get_symm_group(self, mode: str):
        """NVSwitch symmetric-memory context of an axis (created on first use; collective over
        the members of that group)."""
        if mode not in self._symm:
            from ..ops.symm import SymmGroup
            self._symm[mode] = SymmGroup(self.get_group(mode) if mode != "global" else None)
        return self._symm[mode]


torch_parallel_context = ProcessTopology()
tpc = torch_parallel_context


def is_using_pp() -> bool:
    return torch_parallel_context.is_mode_inited("pipe")


def test_comm(verbose: bool = True) -> bool:
    """Communication smoke test over every initialised group (reference: process_topo.py:267-316)
    -- runs on the current accelerator, or on CPU with gloo.  When the native extension and
    symmetric memory are available the custom NVLS all-reduce is exercised as well."""
    ctx = torch_parallel_context
    dev = torch.device("cuda", torch.cuda.current_device()) if (
        torch.cuda.is_available() and dist.get_backend() != "gloo") else torch.device("cpu")
    rank, world = dist.get_rank(), dist.get_world_size()

    def log(msg):
        if verbose and rank == 0:
            print(msg, flush=True)

    x = torch.full((100, 1024), float(rank + 1), device=dev)
    dist.all_reduce(x)
    assert torch.allclose(x, torch.full_like(x, world * (world + 1) / 2))
    log("passed: all_reduce(global)")

    if world > 1:
        buf = torch.full((100, 1024), float(rank), device=dev)
        for src in range(1, world):
            if rank == src:
                dist.send(buf, 0)
            elif rank == 0:
                tmp = torch.empty_like(buf)
                dist.recv(tmp, src)
                assert float(tmp[0, 0]) == float(src)
        dist.barrier()
        log("passed: send/recv to rank 0")

    for mode in ("data", "tensor", "pipe", "model", "moe_dp", "moe_ep", "node"):
        if ctx.is_mode_inited(mode):
            ranks = ctx.get_ranks_in_group(mode)
            y = torch.full((100, 1024), float(rank), device=dev)
            dist.all_reduce(y, group=ctx.get_group(mode))
            assert torch.allclose(y, torch.full_like(y, float(sum(ranks))))
            log(f"passed: all_reduce({mode})")

    if ctx.is_mode_inited("model"):
        t = torch.tensor([10 if ctx.is_first_in_group("model") else 0], dtype=torch.long, device=dev)
        dist.broadcast(t, ctx.get_ranks_in_group("model")[0], group=ctx.get_group("model"))
        assert int(t) == 10
        log("passed: broadcast(model)")

    if ctx.is_mode_inited("tensor"):
        n = ctx.get_group_size("tensor")
        mine = torch.full((16,), float(rank), device=dev)
        outs = [torch.empty_like(mine) for _ in range(n)]
        dist.all_gather(outs, mine, group=ctx.get_group("tensor"))
        assert [float(o[0]) for o in outs] == [float(r) for r in ctx.get_ranks_in_group("tensor")]
        log("passed: all_gather(tensor)")

    if ctx.is_mode_inited("pipe"):
        t = torch.full((8,), float(rank), device=dev)
        if ctx.is_first_in_pipeline_group():
            dist.send(t, ctx.get_next_global_rank("pipe"))
        elif ctx.get_group_rank("pipe") == 1:
            dist.recv(t, ctx.get_prev_global_rank("pipe"))
            assert float(t[0]) == float(ctx.get_prev_global_rank("pipe"))
        log("passed: p2p(pipe)")
    # the package's own NVSwitch data plane: symmetric-memory all-reduce (NVLS multimem kernels or
    # their P2P variant) on every group that lives inside one NVLink domain
    if dev.type == "cuda":
        from ..ops._loader import native
        from ..ops.symm import get_symm_group
        if native() is not None:
            seen = set()
            for mode in ("data", "tensor", "moe_ep", "moe_dp", "node"):
                if not ctx.is_mode_inited(mode) or ctx.get_group_size(mode) < 2:
                    continue
                grp = ctx.get_group(mode)
                if id(grp) in seen:
                    continue
                seen.add(id(grp))
                sg = get_symm_group(grp)
                if not sg.enabled:
                    log(f"skipped: symmetric all_reduce({mode}): {sg.reason}")
                    continue
                buf = sg.alloc(1 << 20)
                ranks = ctx.get_ranks_in_group(mode)
                for n in (1024, 1 << 18):               # one-shot (latency) and two-shot paths
                    v = buf.view(0, (n,), torch.bfloat16)
                    v.fill_(float(rank % 7))
                    buf.all_reduce_(0, n, torch.bfloat16, 1.0)
                    torch.cuda.synchronize()
                    want = float(sum(r % 7 for r in ranks))
                    assert torch.all(v.float() == want), (mode, n, float(v[0]), want)
                log(f"passed: symmetric all_reduce({mode}) "
                    f"[{'NVLS multimem' if buf.has_multicast else 'P2P'}]")
    dist.barrier()
    log("Finished test_comm")
    return True
