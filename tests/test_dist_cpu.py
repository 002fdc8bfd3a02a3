"""Multi-process CPU (gloo) tests of the distributed plumbing: bootstrap, topology, DDP, ZeRO,
sharded EMA, pipeline 1F1B, tensor/sequence parallel layers, grad clipping, MoE-DP.

Pattern follows the reference's example tests (examples/test_ddp.py, test_zero_optim.py,
test_shard_ema.py, model_parallel/*.py): numerical equivalence against a trusted single-process
/ torch implementation -- but with per-rank distinct data, on CPU, under pytest.
"""
import copy

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn

from _mp import run_distributed


class TinyMLP(nn.Module):   # the reference's test_ddp.py MyModule: 10 -> 10 -> 1
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(10, 10)
        self.fc2 = nn.Linear(10, 1)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


# ------------------------------------------------------------------ bootstrap + topology
def _w_topology(rank, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.dist.process_topo import compute_layout
    tpc = tdp.tpc
    cfg = [("data", 2), ("tensor", 2)]
    tpc.setup_process_groups(cfg)
    lay = compute_layout(world, cfg)
    assert tpc.get_ranks_in_group("tensor") == next(g for g in lay["tensor"] if rank in g)
    assert tpc.get_ranks_in_group("data") == next(g for g in lay["data"] if rank in g)
    assert tpc.get_group_size("model") == 2
    assert tpc.get_tp_rank() == rank % 2 and tpc.get_dp_rank() == rank // 2
    assert tpc.is_mode_inited("tensor") and not tpc.is_mode_inited("pipe")
    assert not tdp.is_using_pp()
    assert tpc.get_next_global_rank("tensor") == tpc.get_ranks_in_group("tensor")[(rank % 2 + 1) % 2]
    assert tpc.is_first_group("tensor") == (rank in lay["tensor"][0])
    assert tpc.all_ranks("data") == lay["data"]
    assert tdp.test_comm(verbose=False)
    assert tdp.setup_node_groups(2) is not None       # 4 ranks, 2 per "node"
    assert tdp.setup_node_groups(8) is None
    assert tdp.get_mp_ckpt_suffix() == f"_tp_{rank % 2}.pth"
    # the collective micro-benchmark, every mode, world group and a sub-group (on gloo the NCCL arm)
    from torchdistpackage_b200.dist.py_comm_test import test_collection, test_all2all_balanced
    for mode in ("all_reduce", "all_gather", "reduce_scatter"):
        bw, sec = test_collection(1 << 14, mode, verbose=False, warmup=1, iters=2)   # reference-style unpack
        assert bw >= 0 and sec > 0
    res = test_collection(1 << 14, "all_reduce", tpc.get_group("tensor"), verbose=False, warmup=1, iters=2)
    assert res["world"] == 2 and res["bytes"] == (1 << 14) * 4 and res["busbw_gbs"] > 0
    assert test_all2all_balanced(1 << 12, verbose=False, warmup=1, iters=2)["mode"] == "all_to_all"


def test_topology_and_comm():
    run_distributed(_w_topology, 4)


def _w_moe_groups(rank, world):
    import torchdistpackage_b200 as tdp
    tdp.tpc.setup_process_groups([("data", 4)])
    tdp.tpc.build_moe_groups(moe_ep_size=2)
    assert tdp.tpc.get_ranks_in_group("moe_ep") == ([0, 1] if rank < 2 else [2, 3])
    assert tdp.tpc.get_ranks_in_group("moe_dp") == [rank % 2, rank % 2 + 2]


def test_moe_groups():
    run_distributed(_w_moe_groups, 4)


# ------------------------------------------------------------------ DDP (BASELINE config #1)
def _w_ddp(rank, world, as_view, sync, set_to_none):
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    model = TinyMLP()
    ref = copy.deepcopy(model)
    ddp = tdp.NaiveDDP(model, sync=sync, gradient_as_bucket_view=as_view, bucket_cap_mb=1e-4)
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-2)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    for it in range(5):
        torch.manual_seed(100 * it + rank)           # every rank sees different data
        x = torch.randn(3, 10)
        ddp(x).sum().backward()
        ddp.reduce_gradients()
        # reference: average of all ranks' grads computed by brute force
        xs = []
        for r in range(world):
            torch.manual_seed(100 * it + r)
            xs.append(torch.randn(3, 10))
        ref_opt.zero_grad()
        (sum(ref(xx).sum() for xx in xs) / world).backward()
        for (n, p), (_, q) in zip(ddp.module.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (as_view, sync, set_to_none, it, n)
        opt.step(); ref_opt.step()
        opt.zero_grad(set_to_none=set_to_none)
        for p, q in zip(ddp.module.parameters(), ref.parameters()):
            assert torch.allclose(p, q, atol=1e-6)


def _w_ddp_all(rank, world):
    for as_view, sync, set_to_none in [(True, False, False), (True, False, True),
                                       (False, False, True), (True, True, False)]:
        _w_ddp(rank, world, as_view, sync, set_to_none)
        dist.barrier()


def test_naive_ddp_matches_reference():
    """bucket views / copies, overlapped / synchronous reduction, zero_grad with and without
    set_to_none -- per-rank-distinct data, compared with brute-force averaging every step."""
    run_distributed(_w_ddp_all, 2, timeout=480.0)


class _Branchy(nn.Module):
    """Mixed dtypes, a frozen parameter, a branch that some iterations do not use."""

    def __init__(self):
        super().__init__()
        self.a = nn.Linear(10, 12)
        self.b = nn.Linear(12, 12).double()             # second dtype -> its own buckets
        self.skip = nn.Linear(12, 12)                   # used only when use_skip
        self.frozen = nn.Linear(12, 12)
        for q in self.frozen.parameters():
            q.requires_grad_(False)
        self.out = nn.Linear(12, 3)

    def forward(self, x, use_skip):
        h = torch.tanh(self.a(x))
        h = self.b(h.double()).float() + self.frozen(h)
        if use_skip:
            h = h + self.skip(h)
        return self.out(h)


def _w_ddp_edge_cases(rank, world, as_view):
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    model = _Branchy()
    ref = copy.deepcopy(model)
    # per-parameter groups (every rank creates every group: new_group is collective); all ranks
    # must route the same parameters to a sub-group so that the bucket layouts agree
    sub01, sub2 = dist.new_group([0, 1]), dist.new_group([2]) if world == 3 else None

    class PerParamGroup(tdp.NaiveDDP):                  # the reference's override hook (:95-96)
        def _get_group(self, name, param):
            if name.startswith("out."):
                return sub01 if rank in (0, 1) else sub2
            return self.group

    cls = PerParamGroup if world == 3 else tdp.NaiveDDP
    ddp = cls(model, gradient_as_bucket_view=as_view, bucket_cap_mb=2e-3)
    assert not any(n.startswith("frozen") for n in ddp.reducer.param_bucket)
    assert len({b.dtype for b in ddp.buckets}) == 2 and len(ddp.buckets) >= 3
    for it in range(4):
        use_skip = it % 2 == 0                          # odd iterations leave `skip` without grads
        xs = []
        for r in range(world):
            torch.manual_seed(50 * it + r)
            xs.append(torch.randn(4, 10))
        ddp.zero_grad()
        ddp(xs[rank], use_skip).sum().backward()
        ddp.reduce_gradients()
        ref.zero_grad()
        (sum(ref(x, use_skip).sum() for x in xs) / world).backward()
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if not p.requires_grad:
                assert p.grad is None
            elif n.startswith("out.") and world == 3:
                # reduced over ranks {0, 1} only (rank 2 keeps its local gradient)
                members = [0, 1] if rank in (0, 1) else [2]
                ref2 = copy.deepcopy(ref)
                ref2.zero_grad()
                (sum(ref2(xs[m], use_skip).sum() for m in members) / len(members)).backward()
                want = dict(ref2.named_parameters())[n].grad
                assert torch.allclose(p.grad, want, atol=1e-5), (it, n)
            elif n.startswith("skip.") and not use_skip:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (it, n)
            else:
                assert torch.allclose(p.grad.to(q.grad.dtype), q.grad, atol=1e-5), (it, n)


@pytest.mark.parametrize("world,as_view", [(2, True), (2, False), (3, True)])
def test_naive_ddp_unused_frozen_mixed_dtype_and_per_param_groups(world, as_view):
    run_distributed(_w_ddp_edge_cases, world, as_view)


def _w_ddp_grad_acc(rank, world):
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    model = TinyMLP()
    ref = copy.deepcopy(model)
    ddp = tdp.NaiveDDP(model, gradient_as_bucket_view=True, num_grad_acc_iter=3, reduce_op="sum")
    data = {}
    for mb in range(3):
        for r in range(world):
            torch.manual_seed(10 * mb + r)
            data[(mb, r)] = torch.randn(2, 10)
    for mb in range(3):
        ddp(data[(mb, rank)]).sum().backward()
    ddp.reduce_gradients()
    sum(ref(data[k]).sum() for k in data).backward()
    for p, q in zip(ddp.module.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5)


def test_naive_ddp_grad_accumulation_and_sum():
    run_distributed(_w_ddp_grad_acc, 2)


def _w_ddp_ignore_and_broadcast(rank, world):
    import torchdistpackage_b200 as tdp
    torch.manual_seed(rank)                 # ranks start different: ctor must broadcast rank 0
    model = TinyMLP()
    model._ddp_params_and_buffers_to_ignore = ["fc2.bias"]
    before = model.fc2.bias.detach().clone()
    ddp = tdp.NaiveDDP(model)
    ws = [torch.empty_like(model.fc1.weight) for _ in range(world)]
    dist.all_gather(ws, model.fc1.weight.detach())
    assert all(torch.equal(w, ws[0]) for w in ws)
    assert torch.equal(model.fc2.bias, before)            # ignored: untouched
    assert "fc2.bias" not in ddp.reducer.param_bucket


def test_naive_ddp_broadcast_and_ignore():
    run_distributed(_w_ddp_ignore_and_broadcast, 2)


# ------------------------------------------------------------------ MoE-DP
def _w_moe_dp(rank, world):
    import torchdistpackage_b200 as tdp
    tdp.tpc.setup_process_groups([("data", 4)])
    tdp.tpc.build_moe_groups(moe_dp_size=2)
    torch.manual_seed(rank % 2)             # experts differ across EP ranks, replicated in moe_dp
    expert = nn.Linear(4, 4)
    params = dict(expert.named_parameters())
    tdp.create_moe_dp_hooks(params, tdp.tpc.get_group("moe_dp"),
                            tdp.tpc.get_ranks_in_group("moe_dp")[0])
    torch.manual_seed(50 + rank)
    x = torch.randn(5, 4)
    expert(x).sum().backward()
    tdp.moe_dp_iter_step()
    # expected: mean over my moe_dp group of the local grads
    ranks = tdp.tpc.get_ranks_in_group("moe_dp")
    exp = 0
    for r in ranks:
        torch.manual_seed(50 + r)
        exp = exp + torch.randn(5, 4).sum(0)
    assert torch.allclose(expert.weight.grad, (exp / len(ranks)).expand(4, 4), atol=1e-6)


def test_moe_dp_hooks():
    run_distributed(_w_moe_dp, 4)


# ------------------------------------------------------------------ ZeRO
def _w_zero(rank, world, bucket_size, overlap):
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    model = nn.Sequential(nn.Linear(16, 32), nn.GELU(), nn.Linear(32, 7))
    ref = copy.deepcopy(model)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    zopt = tdp.Bf16ZeroOptimizer(torch.optim.Adam(model.parameters(), lr=1e-2),
                                 bucket_size=bucket_size, overlap_comm=overlap)
    for it in range(4):
        xs = []
        for r in range(world):
            torch.manual_seed(100 * it + r)
            xs.append(torch.randn(6, 16) + r)
        zopt.zero_grad()
        model(xs[rank]).pow(2).sum().backward()
        zopt.step()
        ref_opt.zero_grad()
        (sum(ref(x).pow(2).sum() for x in xs) / world).backward()
        ref_opt.step()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), it
    sd = zopt.state_dict()
    zopt.load_state_dict(sd)
    assert sd["layout"]["world"] == world


def _w_zero_all(rank, world):
    for bucket_size, overlap in [(5e8, False), (256, True)]:
        _w_zero(rank, world, bucket_size, overlap)
        dist.barrier()


def test_zero_optimizer_matches_adam():
    run_distributed(_w_zero_all, 2)


def _w_zero_variants(rank, world, variant):
    """Options of the reference constructor (zero_optim.py:108-109) and the shapes it is used in:
    bf16 parameters with fp32 master weights, bf16 master weights, stage 1, no bucketing, gradient
    accumulation, a non-Adam inner optimizer, two parameter groups, a world size that does not
    divide anything nicely.  Reference: the same optimizer on one process with the world-averaged
    loss (in fp32; the bf16 variants are compared after rounding the reference to bf16)."""
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    bf16 = variant in ("bf16_fp32_master", "bf16_master")
    model = nn.Sequential(nn.Linear(16, 33), nn.GELU(), nn.Linear(33, 7))
    ref = copy.deepcopy(model)
    if bf16:
        model = model.to(torch.bfloat16)
    kw, acc = {}, 1
    make = lambda ps: torch.optim.Adam(ps, lr=1e-2)
    if variant == "bf16_master":
        kw = dict(bf16_master_weights=True, bucketize=False)
    elif variant == "stage1":
        kw = dict(stage=1)
    elif variant == "no_bucketize":
        kw = dict(bucketize=False, overlap_comm=True)
    elif variant == "grad_acc":
        kw, acc = dict(grad_acc_steps=2, bucket_size=256, overlap_comm=True), 2
    elif variant == "sgd_momentum":
        make = lambda ps: torch.optim.SGD(ps, lr=2e-3, momentum=0.9, weight_decay=1e-2)
    if variant == "two_groups":
        groups = lambda m: [dict(params=[m[0].weight, m[0].bias], lr=1e-2),
                            dict(params=[m[2].weight, m[2].bias], lr=3e-3, weight_decay=0.1)]
        zopt = tdp.Bf16ZeroOptimizer(torch.optim.AdamW(groups(model)), bucket_size=300)
        ref_opt = torch.optim.AdamW(groups(ref))
    else:
        zopt = tdp.Bf16ZeroOptimizer(make(model.parameters()), **kw)
        ref_opt = make(ref.parameters())
    assert len(zopt.param_groups) == len(ref_opt.param_groups)
    for it in range(4):
        ref_opt.zero_grad()
        zopt.zero_grad()
        for a in range(acc):
            xs = []
            for r in range(world):
                torch.manual_seed(1000 * it + 10 * a + r)
                xs.append(torch.randn(6, 16) + r)
            x = xs[rank].to(torch.bfloat16) if bf16 else xs[rank]
            (model(x).float().pow(2).sum() / acc).backward()
            zopt.step()                      # counts micro-batches; acts on the last one
            (sum(ref(x).pow(2).sum() for x in xs) / world / acc).backward()
        ref_opt.step()
        for p, q in zip(model.parameters(), ref.parameters()):
            if bf16:
                # bf16 forward/backward noise: the trajectories agree to bf16 resolution
                tol = 6e-2 if variant == "bf16_master" else 3e-2
                assert (p.float() - q).abs().max() <= tol * q.abs().max() + 1e-2, (variant, it)
            else:
                assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (variant, it)
    # every replica holds the same parameters bit for bit
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    # optimizer state lives only for this rank's shard
    n_total = sum(p.numel() for p in model.parameters())
    n_state = sum(m.numel() for m in zopt.master)
    assert n_state < n_total if world > 1 else n_state >= n_total


_ZERO_VARIANTS = ["bf16_fp32_master", "bf16_master", "stage1", "no_bucketize", "grad_acc",
                  "sgd_momentum", "two_groups"]


def _w_zero_variants_all(rank, world):
    for variant in _ZERO_VARIANTS:
        _w_zero_variants(rank, world, variant)
        dist.barrier()


def test_zero_optimizer_variants():
    run_distributed(_w_zero_variants_all, 3, timeout=480.0)


# ------------------------------------------------------------------ hybrid ZeRO (node-local shards)
def _w_hybrid_zero_all(rank, world):
    for mode, as_view, overlap in _HYBRID_VARIANTS:
        _w_hybrid_zero(rank, world, mode, as_view, overlap)
        dist.barrier()


def _w_hybrid_zero(rank, world, mode, as_view, overlap):
    """4 ranks = 2 "nodes" x 2: ZeRO shards inside the node, the cross-node average comes from
    NaiveDDP (either construction order, world or inter-node group) or from ``outer_group``.
    Every variant must track single-process Adam on the world-averaged loss."""
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    node = tdp.setup_node_groups(num_per_node=2)
    inter = tdp.setup_inter_node_groups(num_per_node=2)
    assert node is not None and inter is not None
    model = nn.Sequential(nn.Linear(16, 32), nn.GELU(), nn.Linear(32, 7))
    ref = copy.deepcopy(model)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    ddp = None
    mk_ddp = lambda grp: tdp.NaiveDDP(model, process_group=grp, gradient_as_bucket_view=as_view,
                                      bucket_cap_mb=0.001)
    mk_zero = lambda **kw: tdp.Bf16ZeroOptimizer(torch.optim.Adam(model.parameters(), lr=1e-2),
                                                 dp_group=node, overlap_comm=overlap,
                                                 bucket_size=256, **kw)
    if mode == "ddp_world_first":
        ddp = mk_ddp(None); zopt = mk_zero()
    elif mode == "ddp_inter_first":
        ddp = mk_ddp(inter); zopt = mk_zero()
    elif mode == "zero_first":
        zopt = mk_zero(); ddp = mk_ddp(None)
    else:
        zopt = mk_zero(outer_group=inter)
    fwd = ddp if ddp is not None else model
    for it in range(4):
        xs = []
        for r in range(world):
            torch.manual_seed(100 * it + r)
            xs.append(torch.randn(6, 16) + r)
        zopt.zero_grad()
        fwd(xs[rank]).pow(2).sum().backward()
        if ddp is not None:
            ddp.reduce_gradients()
        zopt.step()
        ref_opt.zero_grad()
        (sum(ref(x).pow(2).sum() for x in xs) / world).backward()
        ref_opt.step()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), (mode, it, (p - q).abs().max())


_HYBRID_VARIANTS = [("ddp_world_first", True, True), ("ddp_world_first", False, False),
                    ("ddp_inter_first", True, True), ("zero_first", True, True),
                    ("outer_group", True, True)]


def test_hybrid_zero_matches_adam():
    """Five compositions, one after the other in the same four processes (a fresh model, fresh
    engines and fresh groups each; the assertion message names the composition)."""
    run_distributed(_w_hybrid_zero_all, 4, timeout=480.0)


# ------------------------------------------------------------------ sharded EMA
def _w_ema(rank, world):
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(0)
    model = nn.Sequential(nn.Linear(8, 16), nn.Linear(16, 16), nn.Linear(16, 3))
    ema = tdp.ShardedEMA(model)
    full = {n: p.detach().clone() for n, p in model.named_parameters()}
    owned = sum(len(part) for part in ema.all_parts)
    assert owned == len(full) and 0 < len(ema.state_dict_shard()) < len(full)
    for it in range(20):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        ema.update(model, decay=0.9)
        for n, p in model.named_parameters():
            full[n].mul_(0.9).add_(p.detach(), alpha=0.1)
    assert ema.verify_with_gt(full)
    sd = ema.state_dict_cpu()
    if rank == 0:
        assert list(sd.keys()) == list(full.keys())
        for n in full:
            assert torch.allclose(sd[n], full[n], atol=1e-6)
    else:
        assert sd is None


def test_sharded_ema():
    run_distributed(_w_ema, 2)


# ------------------------------------------------------------------ pipeline 1F1B
def _w_pipeline(rank, world, pp, n_micro):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import forward_backward, forward_eval, partition_uniform
    tdp.tpc.setup_process_groups([("data", world // pp), ("pipe", pp)])
    tdp.fix_rand(0)
    layers = [nn.Linear(10, 10) for _ in range(6)]
    full = nn.Sequential(*copy.deepcopy(layers))
    stage = nn.Sequential(*partition_uniform(layers))
    pp_rank = tdp.tpc.get_pp_rank()
    dp_rank = tdp.tpc.get_dp_rank()
    torch.manual_seed(7 + dp_rank)
    x = torch.randn(8, 10)
    y = torch.randn(8, 10)
    first, last = tdp.tpc.is_first_in_pipeline_group(), tdp.tpc.is_last_in_pipeline_group()
    mbs = 8 // n_micro

    def fwd(inp):
        if last:
            act, tgt = inp if isinstance(inp, (list, tuple)) else (inp, None)
            if pp == 1:
                act, tgt = inp[0], inp[1]
            return (stage(act) - tgt).pow(2).sum() / 8
        return stage(inp)

    inputs = []
    if first:
        inputs.append(x)
    if last:
        inputs.append(y)
    opt = torch.optim.SGD(stage.parameters(), lr=0.1)
    out = forward_backward(opt, fwd, None, inputs if inputs else None, num_microbatches=n_micro,
                           dtype=torch.float32)
    # oracle: full model on the whole mini-batch
    loss = (full(x) - y).pow(2).sum() / 8
    loss.backward()
    sizes = [len(layers) // pp] * pp
    beg = sum(sizes[:pp_rank])
    for i, lyr in enumerate(stage):
        assert torch.allclose(lyr.weight.grad, full[beg + i].weight.grad, atol=1e-5), (pp_rank, i)
    # eval path
    with torch.no_grad():
        def fwd_eval(inp):
            return stage(inp)
        o = forward_eval(fwd_eval, x if first else None, dtype=torch.float32)
        if last:
            assert torch.allclose(o, full(x), atol=1e-5)


@pytest.mark.parametrize("pp,n_micro", [(2, 4), (3, 4), (3, 2)])
def test_pipeline_1f1b_matches_serial(pp, n_micro):
    run_distributed(_w_pipeline, pp, pp, n_micro)


def _w_pipeline_scatter_gather_and_scaler(rank, world):
    """pipe=2 x tensor=2: stage boundaries send only 1/tp of every activation / gradient and
    all-gather it over the tensor group at the receiver (scatter_gather_tensors=True, reference
    comm.py:108-155); forward_only=True runs the schedule without backward; NativeScalerPP drives
    the loss scale around the schedule; partition_balanced splits by parameter count."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import forward_backward, NativeScalerPP, clip_grad_norm_
    from torchdistpackage_b200.parallel.pipeline_parallel.pipeline_helper import partition_balanced
    tdp.tpc.setup_process_groups([("pipe", 2), ("tensor", 2)])
    tdp.fix_rand(0)
    widths = [(12, 12), (12, 12), (12, 48), (48, 12)]           # unbalanced on purpose
    layers = [nn.Linear(i, o) for i, o in widths]
    full = nn.Sequential(*copy.deepcopy(layers))
    mine = partition_balanced(layers)
    pp_rank = tdp.tpc.get_pp_rank()
    # balanced by parameters: 156+156+624 | 588  beats the uniform 2 | 2 split (312 | 1212)
    assert len(mine) == (3 if pp_rank == 0 else 1), len(mine)
    from torchdistpackage_b200.parallel.pipeline_parallel.pipeline_helper import flat_and_partition
    nested = [nn.Linear(4, 4), [nn.ReLU(), nn.Linear(4, 4)], nn.Identity()]
    assert len(flat_and_partition(nested, flat_level=1)) == 2            # 4 flat items over 2 stages
    stage = nn.Sequential(*mine)
    first, last = tdp.tpc.is_first_in_pipeline_group(), tdp.tpc.is_last_in_pipeline_group()
    torch.manual_seed(3)
    x, y = torch.randn(8, 12), torch.randn(8, 12)

    def fwd(inp):
        if last:
            act, tgt = inp
            return (stage(act) - tgt).pow(2).sum() / 8
        return stage(inp)

    inputs = [x] if first else [y]
    opt = torch.optim.SGD(stage.parameters(), lr=0.1)
    out = forward_backward(opt, fwd, None, inputs, num_microbatches=4, dtype=torch.float32,
                           scatter_gather_tensors=True)
    (full(x) - y).pow(2).sum().div(8).backward()
    beg = 0 if pp_rank == 0 else 3
    for i, lyr in enumerate(stage):
        assert torch.allclose(lyr.weight.grad, full[beg + i].weight.grad, atol=1e-5), (pp_rank, i)
    grads = [lyr.weight.grad.clone() for lyr in stage]

    # forward_only: same loss on the last stage, no gradient touched
    out2 = forward_backward(None, fwd, None, inputs, num_microbatches=4, forward_only=True,
                            dtype=torch.float32, scatter_gather_tensors=True)
    if last:
        mb = slice(6, 8)                                          # the last micro-batch
        want = (full(x[mb]) - y[mb]).pow(2).sum() / 8
        assert torch.allclose(out2.detach(), want.detach(), atol=1e-5)
    for g, lyr in zip(grads, stage):
        assert torch.equal(g, lyr.weight.grad)

    # AMP scaler front-end: backward already done by the schedule -> backward=False
    scaler = NativeScalerPP(enabled=False)
    before = [q.detach().clone() for q in stage.parameters()]
    norm = scaler(out if last else None, opt, clip_grad=0.5, parameters=list(stage.parameters()),
                  backward=False)
    total = torch.sqrt(sum(q.grad.pow(2).sum() for q in full.parameters()))
    assert torch.allclose(norm, total, rtol=1e-4), (float(norm), float(total))    # over both stages
    coef = min(1.0, 0.5 / (float(total) + 1e-6))
    for q0, q, g in zip(before, stage.parameters(), [t for lyr in stage for t in (lyr.weight.grad, lyr.bias.grad)]):
        assert torch.allclose(q, q0 - 0.1 * g, atol=1e-6)          # grads were clipped in place
    for g0, lyr in zip(grads, stage):
        assert torch.allclose(lyr.weight.grad, g0 * coef, atol=1e-6)
    assert isinstance(scaler.state_dict(), dict)

    # enabled scaler, whole call (backward inside): an overflow on ONE stage must make every
    # stage skip the step (found-inf is agreed on over the pipe and tensor groups)
    scaler = NativeScalerPP(enabled=True, init_scale=4.0)
    lin = nn.Linear(4, 4)
    opt2 = torch.optim.SGD(lin.parameters(), lr=0.1)
    w0 = lin.weight.detach().clone()
    loss = lin(torch.ones(2, 4)).sum()
    if rank == 3:
        loss = loss * float("inf")
    scaler(loss, opt2, clip_grad=1.0, parameters=list(lin.parameters()))
    assert torch.equal(lin.weight.detach(), w0), "a stage stepped although another one overflowed"
    assert scaler.state_dict()["scale"] < 4.0                      # backed off everywhere
    opt2.zero_grad()
    scaler(lin(torch.ones(2, 4)).sum(), opt2, clip_grad=1.0, parameters=list(lin.parameters()))
    assert not torch.equal(lin.weight.detach(), w0)


def test_pipeline_scatter_gather_forward_only_and_scaler():
    run_distributed(_w_pipeline_scatter_gather_and_scaler, 4)


def _w_pipeline_static_shapes(rank, world):
    """static_shapes=True: the shape handshake happens on the first call only; later calls must
    give the same results without any metadata message."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import forward_backward
    from torchdistpackage_b200.parallel.pipeline_parallel import comm, pipeline_sched
    tdp.tpc.setup_process_groups([("pipe", world)])
    tdp.fix_rand(0)
    layers = [nn.Linear(8, 8) for _ in range(world)]
    mine = layers[rank]
    first, last = rank == 0, rank == world - 1
    x = torch.randn(8, 8)

    def fwd(inp):
        h = mine(inp[0] if isinstance(inp, (list, tuple)) else inp)
        return h.pow(2).mean() if last else h

    calls = {"n": 0}
    orig = comm.send_obj_meta

    def counting(obj, need_meta=True, next_rank=None):
        if need_meta:
            calls["n"] += 1
        return orig(obj, need_meta, next_rank)
    comm.send_obj_meta = counting
    pipeline_sched.clear_shape_cache()
    outs = []
    for it in range(3):
        mine.zero_grad()
        out = forward_backward(None, fwd, None, x if first else None, num_microbatches=4,
                               dtype=torch.float32, static_shapes=True)
        outs.append(float(out.detach()) if last else 0.0)
    comm.send_obj_meta = orig
    assert calls["n"] == (0 if last else 1), calls          # handshake only on the first call
    if last:
        assert abs(outs[0] - outs[1]) < 1e-7 and abs(outs[1] - outs[2]) < 1e-7


def test_pipeline_static_shapes_skip_the_handshake():
    run_distributed(_w_pipeline_static_shapes, 3)


def test_pipeline_with_data_parallel():
    run_distributed(_w_pipeline, 4, 2, 2)


# ------------------------------------------------------------------ tensor / sequence parallel
def _w_tp_block(rank, world, sequence_parallel):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import Block, ParallelBlock
    from torchdistpackage_b200.parallel.tensor_parallel.transformer import \
        allreduce_sequence_parallel_grads
    tdp.fix_rand(0)
    dim, heads, B, N = 32, 4, 4, 6
    serial = Block(dim, num_heads=heads)
    with torch.no_grad():
        for n, p in serial.named_parameters():
            if p.dim() == 2:
                p.mul_(0.2).sub_(0.1)
            elif "bias" in n:
                p.copy_(torch.rand_like(p) - 0.5)
    par = ParallelBlock(dim, num_heads=heads, sequence_parallel=sequence_parallel)
    par.init_from_full(serial)
    x = torch.randn(B, N, dim)
    g = torch.randn(B, N, dim)
    xs = x.clone().requires_grad_(True)
    ys = serial(xs); ys.backward(g)
    xp = x.clone().requires_grad_(True)
    yp = par(xp)
    if sequence_parallel:
        k = B // world
        assert torch.allclose(yp, ys[rank * k:(rank + 1) * k], atol=1e-4)
        yp.backward(g[rank * k:(rank + 1) * k])
        allreduce_sequence_parallel_grads(par)
        assert torch.allclose(xp.grad[rank * k:(rank + 1) * k], xs.grad[rank * k:(rank + 1) * k], atol=1e-4)
    else:
        assert torch.allclose(yp, ys, atol=1e-4)
        yp.backward(g)
        assert torch.allclose(xp.grad, xs.grad, atol=1e-4)    # input grad is correct in plain TP
    h = 4 * dim // world
    assert torch.allclose(par.mlp.fc1.linear.weight.grad,
                          serial.mlp.fc1.weight.grad[:, rank * h:(rank + 1) * h], atol=1e-4)
    assert torch.allclose(par.mlp.fc2.linear.weight.grad,
                          serial.mlp.fc2.weight.grad[rank * h:(rank + 1) * h], atol=1e-4)
    d = dim // world
    assert torch.allclose(par.attn.proj.linear.weight.grad,
                          serial.attn.proj.weight.grad[rank * d:(rank + 1) * d], atol=1e-4)
    assert torch.allclose(par.ln_1.weight.grad, serial.ln_1.weight.grad, atol=1e-4)
    assert torch.allclose(par.mlp.fc2.linear.bias.grad, serial.mlp.fc2.bias.grad, atol=1e-4)


def _w_tp_block_both(rank, world):
    for sp in (False, True):
        _w_tp_block(rank, world, sp)
        dist.barrier()


def test_tp_block_matches_serial():
    run_distributed(_w_tp_block_both, 2)


def _w_tp_transformer(rank, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel.tensor_parallel.transformer import Transformer
    tdp.fix_rand(0)
    serial = Transformer(16, num_heads=2, depth=2, tensor_parallel=False, sequence_parallel=False)
    with torch.no_grad():
        for p in serial.parameters():
            if p.dim() == 2:
                p.mul_(0.3).sub_(0.15)
    par = Transformer(16, num_heads=2, depth=2, tensor_parallel=True, sequence_parallel=True)
    for pb, sb in zip(par.blocks, serial.blocks):
        pb.init_from_full(sb)
    x = torch.randn(4, 5, 16)
    out_p, out_s = par(x), serial(x)
    assert torch.allclose(out_p, out_s, atol=1e-4)
    # gradient parity with the serial stack (replicated loss on the gathered output)
    from torchdistpackage_b200.parallel.tensor_parallel.transformer import allreduce_sequence_parallel_grads
    out_p.pow(2).sum().backward()
    out_s.pow(2).sum().backward()
    allreduce_sequence_parallel_grads(par)
    tp = tdp.tpc.get_tp_rank() if tdp.tpc.is_mode_inited("tensor") else rank
    for pb, sb in zip(par.blocks, serial.blocks):
        assert torch.allclose(pb.ln_1.weight.grad, sb.ln_1.weight.grad, rtol=1e-3, atol=1e-4)
        assert torch.allclose(pb.ln_2.bias.grad, sb.ln_2.bias.grad, rtol=1e-3, atol=1e-4)
        g_full = sb.mlp.fc2.weight.grad            # row-parallel: split along the input features
        g_mine = pb.mlp.fc2.linear.weight.grad
        n = g_mine.shape[0]
        assert torch.allclose(g_mine, g_full[rank * n:(rank + 1) * n], rtol=1e-3, atol=1e-4)


def test_tp_sp_transformer_forward():
    run_distributed(_w_tp_transformer, 2)


# ------------------------------------------------------------------ gradient clipping
def _w_clip(rank, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import clip_grad_norm_
    tdp.tpc.setup_process_groups([("pipe", 2), ("tensor", 2)])
    torch.manual_seed(0)
    full = [torch.randn(6, 4), torch.randn(8)]      # stage 0: a TP-sharded matrix; + replicated vec
    stage = tdp.tpc.get_pp_rank()
    tp = tdp.tpc.get_tp_rank()
    torch.manual_seed(10 + stage)
    w_full = torch.randn(6, 4)
    b_full = torch.randn(8)
    w = nn.Parameter(torch.zeros(3, 4)); w.grad = w_full[tp * 3:(tp + 1) * 3].clone()
    w.tensor_model_parallel = True
    b = nn.Parameter(torch.zeros(8)); b.grad = b_full.clone()
    norm = clip_grad_norm_([w, b], max_norm=1.0)
    expect_sq = 0.0
    for s in range(2):
        torch.manual_seed(10 + s)
        expect_sq += torch.randn(6, 4).pow(2).sum() + torch.randn(8).pow(2).sum()
    assert torch.allclose(norm, expect_sq.sqrt(), rtol=1e-5)
    coef = 1.0 / (expect_sq.sqrt() + 1e-6)
    assert torch.allclose(b.grad, b_full * coef, rtol=1e-5)


def test_clip_grad_norm_model_parallel():
    run_distributed(_w_clip, 4)


def _w_clip_zero_tp(rank, world):
    """ZeRO (over the data group) x tensor parallel: TP-replicated parameters must be counted
    once in the global norm, TP shards on every TP rank."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import clip_grad_norm_
    tdp.tpc.setup_process_groups([("data", 2), ("tensor", 2)])
    tp, dp = tdp.tpc.get_tp_rank(), tdp.tpc.get_group_rank("data")
    torch.manual_seed(0)
    w_full, b_full = torch.randn(6, 40), torch.randn(72)      # same gradient on both DP replicas
    w = nn.Parameter(torch.zeros(3, 40)); w.tensor_model_parallel = True
    b = nn.Parameter(torch.zeros(72))
    zopt = tdp.Bf16ZeroOptimizer(torch.optim.SGD([w, b], lr=0.1), dp_group=tdp.tpc.get_group("data"),
                                 bucket_size=64)
    with torch.no_grad():
        w.grad.copy_(w_full[tp * 3:(tp + 1) * 3])
        b.grad.copy_(b_full)
    norm = clip_grad_norm_([], max_norm=1.0, zero_optimizer=zopt)
    expect = (w_full.pow(2).sum() + b_full.pow(2).sum()).sqrt()
    assert torch.allclose(norm, expect, rtol=1e-5), (norm, expect)


def test_clip_grad_norm_zero_with_tensor_parallel():
    run_distributed(_w_clip_zero_tp, 4)


# ------------------------------------------------------------------ MoE layer (expert parallel)
def _w_moe_layer(rank, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.moe import MoELayer
    tdp.tpc.setup_process_groups([("data", world)])
    tdp.tpc.build_moe_groups(moe_ep_size=2)
    ep_group = tdp.tpc.get_group("moe_ep")
    ep_rank = tdp.tpc.get_group_rank("moe_ep")
    dim, hid, E = 16, 32, 4
    torch.manual_seed(0)
    serial = MoELayer(dim, hid, num_experts=E, top_k=2, capacity_factor=8.0, expert_parallel=False)
    par = MoELayer(dim, hid, num_experts=E, top_k=2, capacity_factor=8.0, ep_group=ep_group)
    with torch.no_grad():
        par.gate.wg.copy_(serial.gate.wg)
        el = E // 2
        for name in ("w1", "b1", "w2", "b2"):
            getattr(par.experts, name).copy_(getattr(serial.experts, name)[ep_rank * el:(ep_rank + 1) * el])
    torch.manual_seed(100 + ep_rank)
    x = torch.randn(12, dim, requires_grad=True)
    y, aux = par(x)
    (y.pow(2).sum() + aux).backward()
    xs = x.detach().clone().requires_grad_(True)
    ys, auxs = serial(xs)
    (ys.pow(2).sum() + auxs).backward()
    assert torch.allclose(y, ys, atol=1e-5)
    assert torch.allclose(x.grad, xs.grad, atol=1e-4)
    # expert grads: my local experts see tokens from both EP ranks; compare against the serial
    # layer fed with the concatenation of both ranks' tokens
    xs_all = []
    for r in range(2):
        torch.manual_seed(100 + r)
        xs_all.append(torch.randn(12, dim))
    serial.zero_grad()
    tot = 0
    for xa in xs_all:
        ya, aa = serial(xa)
        tot = tot + ya.pow(2).sum() + aa
    tot.backward()
    assert torch.allclose(par.experts.w1.grad, serial.experts.w1.grad[ep_rank * el:(ep_rank + 1) * el], atol=1e-4)


def test_moe_layer_expert_parallel_matches_serial():
    run_distributed(_w_moe_layer, 2)


# ------------------------------------------------------------------ full mixed parallel (config #5)
def _w_mixed_parallel_gpt(rank, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.models.gpt2 import GPT2Config
    from torchdistpackage_b200.models.gpt2_parallel import GPT2PipelineStage
    from torchdistpackage_b200.parallel import forward_backward
    tdp.tpc.setup_process_groups([("data", world // 4), ("pipe", 2), ("tensor", 2)])
    cfg = GPT2Config(vocab_size=128, n_layer=4, n_head=4, d_model=64, seq_len=32)
    tdp.fix_rand(0)
    stage = GPT2PipelineStage(cfg, tp_group=tdp.tpc.get_group("tensor"), sequence_parallel=True)
    opt = tdp.Bf16ZeroOptimizer(torch.optim.AdamW(stage.parameters(), lr=1e-2),
                                dp_group=tdp.tpc.get_group("data"), overlap_comm=False)
    torch.manual_seed(5 + tdp.tpc.get_dp_rank())
    tok = torch.randint(0, 128, (8, 33))
    losses = []
    for it in range(6):
        out = forward_backward(opt, stage.forward_fn(), None,
                               stage.stage_inputs(tok[:, :-1], tok[:, 1:]), num_microbatches=2,
                               dtype=torch.float32)
        stage.allreduce_replicated_grads()
        opt.step()
        if stage.last:
            losses.append(float(out))
    if stage.last:
        assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    # tensor-parallel replicas of replicated parameters must stay identical
    if stage.last:
        w = stage.ln_f.weight.detach().clone()
        ws = [torch.empty_like(w) for _ in range(2)]
        dist.all_gather(ws, w, group=tdp.tpc.get_group("tensor"))
        assert torch.allclose(ws[0], ws[1], atol=1e-6)


def test_mixed_parallel_gpt_dp_pp_tp_zero():
    run_distributed(_w_mixed_parallel_gpt, 8)


def _fd_share_worker(rank, world):
    import os
    import torch.distributed as dist
    from torchdistpackage_b200.ops.symm import _share_fd
    r, w = os.pipe()
    os.write(w, f"rank{rank}".encode())
    got = _share_fd(dist.group.WORLD, r, list(range(world)))
    assert sorted(got) == [p for p in range(world) if p != rank]
    dist.barrier()
    # a pipe read end is shared, so each message is consumed exactly once: everyone reads from
    # the pipe of the next rank only
    nxt = (rank + 1) % world
    assert os.read(got[nxt], 16) == f"rank{nxt}".encode()
    only0 = _share_fd(dist.group.WORLD, r, [0])
    assert (rank == 0 and only0 == {}) or (rank != 0 and list(only0) == [0])


def test_symm_fd_passing():
    """native symmetric-memory backend: POSIX fds travel between ranks with SCM_RIGHTS"""
    run_distributed(_fd_share_worker, 3)


# ------------------------------------------------------------------ DDP + direct weight gradients
def _w_ddp_direct_wgrad(rank, world):
    """Two ranks, ops.linear layers (kernel emulated with torch), weight gradients written straight
    into the NaiveDDP bucket views: buckets must be reduced only when every gradient is final."""
    import torch.nn as nn
    import torch.nn.functional as F
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.ops import linear as L

    def fake_gemm(a, b, *, trans_a=False, trans_b=False, out=None, out_dtype=None, bias=None,
                  residual=None, aux_in=None, aux_out=None, act=0, accumulate=False, alpha=1.0, **kw):
        y = ((a.t() if trans_a else a).double() @ (b.t() if trans_b else b).double()) * alpha
        if bias is not None:
            y = y + bias.double()
        if aux_out is not None:
            aux_out.copy_(y.to(aux_out.dtype))
        if act == L.ACT_GELU_TANH:
            y = F.gelu(y, approximate="tanh")
        if act == L.ACT_DGELU_TANH:
            with torch.enable_grad():
                z = aux_in.double().requires_grad_(True)
                g, = torch.autograd.grad(F.gelu(z, approximate="tanh").sum(), z)
            y = y * g
        if residual is not None:
            y = y + residual.double()
        if out is None:
            return y.to(out_dtype or a.dtype)
        out.add_(y.to(out.dtype)) if accumulate else out.copy_(y.to(out.dtype))
        return out

    L.gemm = fake_gemm
    L._native_ok = lambda *ts: True
    L.colsum = lambda x, out_dtype=None: x.sum(0).to(out_dtype or x.dtype)
    L._FUSED_WGRAD = True

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.w1 = nn.Parameter(torch.randn(8, 16) * 0.2)
            self.b1 = nn.Parameter(torch.zeros(16))
            self.w2 = nn.Parameter(torch.randn(16, 8) * 0.2)
            self.b2 = nn.Parameter(torch.zeros(8))
            self.w_tied = nn.Parameter(torch.randn(8, 8) * 0.2)

        def forward(self, x):
            h = L.mlp(x, self.w1, self.b1, self.w2, self.b2, layout="kn", act="gelu_tanh", residual=x)
            h = L.linear(h, self.w_tied, None, layout="kn")
            return L.linear(h, self.w_tied, None, layout="kn", residual=h).pow(2).mean()

    tdp.fix_rand(0)
    model = Net().double()
    ref = copy.deepcopy(model)
    ddp = tdp.NaiveDDP(model, gradient_as_bucket_view=True, bucket_cap_mb=1e-4, num_grad_acc_iter=2)
    red = ddp.reducer
    for name, p in red.params.items():             # what the reducer does itself on CUDA
        if p.dim() == 2:
            p._tdp_main_grad = red.param_bucket[name].views[name]
            p._tdp_grad_fresh = True
    for step in range(2):
        ddp.zero_grad()
        ref.zero_grad(set_to_none=True)
        for mb in range(2):
            torch.manual_seed(1000 * step + 10 * mb + rank)
            ddp(torch.randn(4, 8, dtype=torch.double)).backward()
        ddp.reduce_gradients()
        for mb in range(2):
            for r in range(world):
                torch.manual_seed(1000 * step + 10 * mb + r)
                (ref(torch.randn(4, 8, dtype=torch.double)) / world).backward()
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-10), (step, n)


def test_naive_ddp_direct_weight_gradients_two_ranks():
    run_distributed(_w_ddp_direct_wgrad, 2)


# ------------------------------------------------------------------ end-to-end example: resume
def test_train_example_checkpoint_resume_reproduces_uninterrupted_run(tmp_path):
    """examples/train_gpt2_ddp.py (NaiveDDP + BucketAdamW + watchdog + metrics + async checkpoint):
    3 steps + resume to 6 gives the same loss trajectory as 6 uninterrupted steps (2 ranks, gloo)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(port, *extra):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
               "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(root, "examples", "train_gpt2_ddp.py"), "--cpu", "--model", "tiny", *extra]
        r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        return [float(x) for x in re.findall(r"^step \d+ loss ([0-9.]+)", r.stdout, flags=re.M)]

    full = run(29681, "--steps", "6", "--out", str(tmp_path / "full"))
    part = run(29682, "--steps", "3", "--ckpt-every", "3", "--out", str(tmp_path / "part"))
    rest = run(29683, "--steps", "6", "--resume", "--out", str(tmp_path / "part"))
    assert len(full) == 6 and part == full[:3] and rest == full[3:], (full, part, rest)
    assert (tmp_path / "part" / "metrics.jsonl").exists()


# ------------------------------------------------------------------ remaining public surface
def _w_topology_queries_and_checkpoint(rank, world, tmp):
    """Every ``tpc`` query of the reference (process_topo.py:147-259) on an 8-rank
    data x pipe x tensor layout; model-parallel checkpoint files written by DP replica 0 only,
    sharded state per DP rank; EMA shard save / load; the bidirectional p2p wrappers."""
    import os
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.dist.model_parallel_ckpt import save_mp_checkpoint, load_mp_checkpoint
    from torchdistpackage_b200.dist.launch import get_cpu_group
    from torchdistpackage_b200.parallel.pipeline_parallel import comm
    from torchdistpackage_b200.parallel.tensor_parallel import tp_utils
    tpc = tdp.tpc
    tpc.setup_process_groups([("data", 2), ("pipe", 2), ("tensor", 2)])
    dp, pp, tp = rank // 4, (rank // 2) % 2, rank % 2
    assert (tpc.get_dp_rank(), tpc.get_pp_rank(), tpc.get_tp_rank()) == (dp, pp, tp)
    assert (tpc.get_dp_size(), tpc.get_pp_size(), tpc.get_tp_size(), tpc.get_mp_size()) == (2, 2, 2, 4)
    assert tpc.get_mp_rank() == rank % 4
    assert tpc.is_first_in_tensor_group() == (tp == 0) and tpc.is_last_in_tensor_group() == (tp == 1)
    assert tpc.is_first_in_data_group() == (dp == 0) and tpc.is_last_in_data_group() == (dp == 1)
    assert tpc.is_first_in_model_group() == (rank % 4 == 0) and tpc.is_last_in_model_group() == (rank % 4 == 3)
    assert tpc.is_first_in_pipeline_group() == (pp == 0) and tpc.is_last_in_pipeline_group() == (pp == 1)
    assert tpc.all_dp_ranks() == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert tpc.get_group("global") is None and tdp.is_using_pp()
    assert tpc.get_prev_global_rank("pipe") == tpc.get_next_global_rank("pipe")     # ring of 2
    assert tpc.setup_node_groups(4) is not None and tpc.get_group_size("node") == 4
    assert get_cpu_group() is dist.group.WORLD                                      # gloo already
    # TP helpers follow the module-level TP group
    tp_utils.set_tp_group(tpc.get_group("tensor"))
    assert tp_utils.get_tensor_model_parallel_world_size() == 2
    x = torch.full((2, 3), float(tp))
    shard = tp_utils.maybe_split_into_sequence_parallel(torch.arange(4.0).view(4, 1))
    assert shard.shape[0] == 2 and tp_utils.is_squence_parallel_tensor(shard)
    assert torch.equal(tp_utils.maybe_gather_from_sequence_parallel(shard), torch.arange(4.0).view(4, 1))
    assert tp_utils.maybe_gather_from_sequence_parallel(x) is x                     # not SP: untouched
    tp_utils.reset_tp_group()

    # checkpoint naming + writers
    prefix = os.path.join(tmp, "ckpt", "model")
    assert tdp.get_mp_ckpt_suffix() == f"_tp_{tp}_pp_{pp}.pth"
    path = save_mp_checkpoint(prefix, {"w": torch.full((2,), float(rank % 4))},
                              sharded_state={"m": torch.full((1,), float(rank))})
    assert path.endswith(f"_tp_{tp}_pp_{pp}.pth")
    files = sorted(os.listdir(os.path.dirname(prefix)))
    assert len([f for f in files if "_shard" not in f]) == 4 and len([f for f in files if "_shard" in f]) == 8
    state, sh = load_mp_checkpoint(prefix, with_shard=True)
    assert float(state["w"][0]) == rank % 4 and float(sh["m"][0]) == rank      # replica 0 wrote the model part

    # EMA shard round trip
    model = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 4))
    ema = tdp.ShardedEMA(model, group=tpc.get_group("data"))
    saved = {n: t.clone() for n, t in ema.state_dict_shard().items()}
    ema.update(model, decay=0.5)
    with torch.no_grad():
        for q in model.parameters():
            q.add_(1.0)
    ema.update(model, decay=0.5)
    assert any(not torch.equal(saved[n], t) for n, t in ema.state_dict_shard().items())
    ema.load_state_dict_shard(saved)
    assert all(torch.equal(saved[n], t) for n, t in ema.state_dict_shard().items())

    # bidirectional wrappers between the two pipeline stages (each is the other's prev AND next)
    a = torch.full((2, 3), float(rank))
    peer = tpc.get_next_global_rank("pipe")
    got = comm.send_forward_recv_forward(a, a.shape, dtype=torch.float32)
    assert torch.equal(got, torch.full((2, 3), float(peer)))
    got = comm.send_backward_recv_backward(a + 1, a.shape, dtype=torch.float32)
    assert torch.equal(got, torch.full((2, 3), float(peer + 1)))


def test_topology_queries_checkpoint_ema_shard_and_p2p_wrappers(tmp_path):
    run_distributed(_w_topology_queries_and_checkpoint, 8, str(tmp_path))


def _w_p2p_ring_exchange(rank, world):
    """send_forward_backward_recv_forward_backward on a ring of 3 stages: one batched exchange
    sends to both neighbours and receives from both (prev / next wrap around, as
    tpc.get_prev/next_global_rank do in the reference, process_topo.py:222-234)."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel.pipeline_parallel import comm
    tdp.tpc.setup_process_groups([("pipe", 3)])
    prev, nxt = tdp.tpc.get_prev_global_rank("pipe"), tdp.tpc.get_next_global_rank("pipe")
    assert (prev, nxt) == ((rank - 1) % 3, (rank + 1) % 3)
    a = torch.full((2, 3), float(rank))
    for _ in range(2):
        t, g = comm.send_forward_backward_recv_forward_backward([a, a * 10], a + 100, [a.shape, a.shape],
                                                                 a.shape, dtype=torch.float32)
        assert torch.equal(t[0], torch.full((2, 3), float(prev))) and torch.equal(t[1], torch.full((2, 3), 10.0 * prev))
        assert torch.equal(g, torch.full((2, 3), float(nxt + 100)))


def test_pipeline_p2p_bidirectional_ring_exchange():
    run_distributed(_w_p2p_ring_exchange, 3)


def _w_ddp_small_api(rank, world):
    """NaiveDDP / MoEDP methods around the main loop: hook removal, re-broadcast, accumulation
    count, manual dispatch, comm sync; Bf16ZeroOptimizer.state passthrough."""
    import torchdistpackage_b200 as tdp
    tdp.fix_rand(rank)                               # replicas start different
    model = TinyMLP()
    ddp = tdp.NaiveDDP(model, gradient_as_bucket_view=True)
    ddp.sync_comm()
    ddp.set_num_grad_acc_iter(2)
    for mb in range(2):
        torch.manual_seed(10 * mb + rank)
        ddp(torch.randn(3, 10)).sum().backward()
    ddp.reduce_gradients()
    g = model.fc1.weight.grad.clone()
    gs = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gs, g)
    assert all(torch.allclose(gs[0], t) for t in gs)             # reduced once, after 2 micro-batches
    with torch.no_grad():
        model.fc1.weight.add_(float(rank))                       # drift apart ...
    ddp.broadcast_params()                                       # ... and re-sync from rank 0
    ws = [torch.empty_like(model.fc1.weight) for _ in range(world)]
    dist.all_gather(ws, model.fc1.weight.detach())
    assert all(torch.equal(ws[0], t) for t in ws)
    ddp.remove_hooks()
    ddp.zero_grad()
    torch.manual_seed(100 + rank)
    model(torch.randn(3, 10)).sum().backward()
    ddp.reduce_gradients()                                       # finalize still reduces what it holds
    assert not hasattr(model.fc1.weight, "_tdp_reducer")

    # MoEDP on a sub-group: manual dispatch of a gradient produced outside autograd
    experts = {"e.w": nn.Parameter(torch.full((4, 4), float(rank)))}
    mo = tdp.create_moe_dp_hooks(experts, None, 0, overlap_comm=True)
    assert torch.equal(experts["e.w"].detach(), torch.zeros(4, 4))      # broadcast from rank 0
    experts["e.w"].grad = None
    (experts["e.w"] * float(rank + 1)).sum().backward()
    tdp.moe_dp_iter_step()
    assert torch.allclose(experts["e.w"].grad, torch.full((4, 4), sum(range(1, world + 1)) / world))
    with torch.no_grad():
        experts["e.w"].add_(float(rank))
    mo.broadcast_params()
    assert torch.equal(experts["e.w"].detach(), torch.zeros(4, 4))
    mo.remove_hooks()

    z = tdp.Bf16ZeroOptimizer(torch.optim.Adam(TinyMLP().parameters(), lr=1e-3))
    assert isinstance(z.state, dict) or hasattr(z.state, "keys")


def test_naive_ddp_and_moe_dp_small_api():
    run_distributed(_w_ddp_small_api, 2)


# ------------------------------------------------------------------ SLURM bootstrap
def _slurm_entry(rank, world, port, err_q):
    import os
    import traceback
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR"):
        os.environ.pop(k, None)
    os.environ.update(SLURM_JOB_ID="4242", SLURM_PROCID=str(rank), SLURM_NTASKS=str(world),
                      SLURM_NODELIST="localhost", SLURM_LOCALID=str(rank), MASTER_PORT=str(port))
    try:
        import torchdistpackage_b200 as tdp
        from torchdistpackage_b200.dist.launch_from_slurm import setup_distributed   # reference path
        r, w, p, addr = setup_distributed("nccl")           # no GPU here: degrades to gloo
        assert (r, w, p, addr) == (rank, world, port, "localhost"), (r, w, p, addr)
        assert os.environ["RANK"] == str(rank) and os.environ["WORLD_SIZE"] == str(world)
        t = torch.tensor([float(rank)])
        dist.all_reduce(t)
        assert float(t) == sum(range(world))
        tdp.shutdown_distributed()
    except Exception:
        err_q.put((rank, traceback.format_exc()))
        raise


def test_setup_distributed_from_slurm_environment():
    """The reference's primary launch mode (launch_from_slurm.py:29-51): rank / world / master come
    from SLURM variables; the tuple it returns is complete (the reference leaves ``addr`` unbound
    under torchrun, :62)."""
    import torch.multiprocessing as mp
    from _mp import _free_port
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_slurm_entry, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    msgs = []
    while not q.empty():
        msgs.append(q.get())
    assert all(p.exitcode == 0 for p in procs) and not msgs, msgs


# ------------------------------------------------------------------ user-side habits the engines must survive
def _w_engine_edge_paths(rank, world):
    """``zero_grad(set_to_none=True)`` (the torch default) under the ZeRO optimizer: the
    gradient views are dropped, autograd creates fresh tensors, the hook re-attaches them.
    Resume from a checkpoint continues the exact trajectory; a checkpoint of another layout is
    refused.  NaiveDDP in copy mode with a parameter that got no gradient."""
    import torchdistpackage_b200 as tdp

    def data(it):
        xs = []
        for r in range(world):
            torch.manual_seed(31 * it + r)
            xs.append(torch.randn(5, 10))
        return xs

    tdp.fix_rand(0)
    model = TinyMLP()
    ref = copy.deepcopy(model)
    zopt = tdp.Bf16ZeroOptimizer(torch.optim.Adam(model.parameters(), lr=1e-2), bucket_size=128,
                                 overlap_comm=True)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    sd = None
    for it in range(6):
        xs = data(it)
        for p in model.parameters():
            p.grad = None                                  # what optimizer.zero_grad() does by default
        for mg in zopt.master_grad:
            mg.zero_()
        model(xs[rank]).sum().backward()
        zopt.step()
        ref_opt.zero_grad()
        (sum(ref(x).sum() for x in xs) / world).backward()
        ref_opt.step()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, atol=1e-6), it
        if it == 2:
            sd = copy.deepcopy(zopt.state_dict())
            snap = [p.detach().clone() for p in model.parameters()]
    final = [p.detach().clone() for p in model.parameters()]

    # resume from step 3 in a fresh optimizer object: same end point
    tdp.fix_rand(1)
    model2 = TinyMLP()
    with torch.no_grad():
        for p, s_ in zip(model2.parameters(), snap):
            p.copy_(s_)
    zopt2 = tdp.Bf16ZeroOptimizer(torch.optim.Adam(model2.parameters(), lr=1e-2), bucket_size=128,
                                  overlap_comm=True)
    zopt2.load_state_dict(sd)
    for it in range(3, 6):
        zopt2.zero_grad(set_to_none=True)
        model2(data(it)[rank]).sum().backward()
        zopt2.step()
    for p, q in zip(model2.parameters(), final):
        assert torch.allclose(p, q, atol=1e-6)
    bad = copy.deepcopy(sd)
    bad["layout"]["world"] = world + 1
    with pytest.raises(ValueError):
        zopt2.load_state_dict(bad)
    with pytest.raises(TypeError):                         # one dtype per optimizer (as the reference)
        mixed = [nn.Parameter(torch.zeros(4)), nn.Parameter(torch.zeros(4, dtype=torch.float64))]
        tdp.Bf16ZeroOptimizer(torch.optim.Adam(mixed, lr=1e-3))

    # NaiveDDP, copy mode (no bucket views), one parameter without gradient on odd steps
    model3 = _Branchy()
    ref3 = copy.deepcopy(model3)
    ddp = tdp.NaiveDDP(model3, gradient_as_bucket_view=False, bucket_cap_mb=1e-3)
    for it in range(3):
        xs = [x[:, :10] for x in data(it)]
        ddp.zero_grad(set_to_none=True)
        ddp(xs[rank], it % 2 == 0).sum().backward()
        ddp.reduce_gradients()
        ref3.zero_grad()
        (sum(ref3(x, it % 2 == 0).sum() for x in xs) / world).backward()
        for (n, p), (_, q) in zip(model3.named_parameters(), ref3.named_parameters()):
            if q.grad is None or (n.startswith("skip.") and it % 2 == 1):
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (it, n)
            else:
                assert torch.allclose(p.grad.to(q.grad.dtype), q.grad, atol=1e-5), (it, n)
    with pytest.raises(ValueError):
        tdp.NaiveDDP(TinyMLP(), reduce_op="max")
    with pytest.raises(TypeError):
        tdp.NaiveDDP(TinyMLP(), no_such_option=1)


def test_engines_survive_set_to_none_resume_and_unused_parameters():
    run_distributed(_w_engine_edge_paths, 2)


def _w_pipeline_multi_tensor_boundary(rank, world):
    """Stages hand over TWO tensors (hidden state + a skip stream), the user supplies bwd_fn
    (reference pipeline_sched.py:36-69; its `cur_inputs.append(*input_obj_from_prev)` only works
    for one tensor, :24), 3 stages, 3 micro-batches."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import forward_backward, forward_eval
    tdp.tpc.setup_process_groups([("pipe", 3)])
    tdp.fix_rand(0)
    layers = [nn.Linear(6, 6) for _ in range(3)]
    full = copy.deepcopy(layers)
    mine = layers[rank]
    first, last = rank == 0, rank == 2
    torch.manual_seed(11)
    x, y = torch.randn(6, 6), torch.randn(6, 6)
    calls = []

    def fwd(inp):
        if first:
            h = mine(inp)
            return [h, inp * 2.0]                            # (hidden, skip)
        if last:
            h, skip, tgt = inp
            return ((mine(h) + skip - tgt) ** 2).sum() / 6
        h, skip = inp
        return [torch.tanh(mine(h)), skip + h]

    def bwd(output, output_grad):
        calls.append(1)
        if output_grad is None:
            output.backward()
        else:
            pairs = [(o, g) for o, g in zip(output, output_grad) if o.requires_grad]
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])

    inputs = [x] if first else ([y] if last else None)
    opt = torch.optim.SGD(mine.parameters(), lr=0.1)
    out = forward_backward(opt, fwd, bwd, inputs, num_microbatches=3, dtype=torch.float32)
    assert len(calls) == 3
    g_user = [mine.weight.grad.clone(), mine.bias.grad.clone()]
    # the built-in backward gives the same gradients (stage 0 forwards plain data as its 2nd output)
    forward_backward(opt, fwd, None, inputs, num_microbatches=3, dtype=torch.float32)
    assert torch.allclose(mine.weight.grad, g_user[0], atol=1e-6)
    assert torch.allclose(mine.bias.grad, g_user[1], atol=1e-6)
    # oracle
    xr = x.clone().requires_grad_(True)
    h0 = full[0](xr)
    h1 = torch.tanh(full[1](h0))
    loss = ((full[2](h1) + (xr * 2.0 + h0) - y) ** 2).sum() / 6
    loss.backward()
    assert torch.allclose(mine.weight.grad, full[rank].weight.grad, atol=1e-5), rank
    assert torch.allclose(mine.bias.grad, full[rank].bias.grad, atol=1e-5), rank
    with torch.no_grad():
        o = forward_eval(fwd, [x] if first else ([y] if last else None), dtype=torch.float32)
        if last:
            assert torch.allclose(o, loss.detach(), atol=1e-5)


def test_pipeline_multi_tensor_stage_boundary_with_user_backward():
    run_distributed(_w_pipeline_multi_tensor_boundary, 3)


def _w_clip_variants(rank, world):
    """clip_grad_norm_ corners: a single tensor, the max norm across model-parallel groups,
    the non-finite check, nothing to clip; q/k/v-aligned weight slicing with a bias."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import clip_grad_norm_, ColParallelLinear
    tdp.tpc.setup_process_groups([("pipe", 2), ("tensor", 2)])
    p = nn.Parameter(torch.zeros(4))
    p.grad = torch.full((4,), float(rank + 1))
    p.is_tp_shard = True
    n = clip_grad_norm_(p, max_norm=100.0, norm_type=float("inf"))
    assert float(n) == 4.0                                    # max over all 4 ranks
    n2 = clip_grad_norm_(p, max_norm=1.0)                     # shards: sum of squares over tp and pp
    assert torch.allclose(n2, torch.tensor(4 * (1 + 4 + 9 + 16.0)).sqrt())
    assert torch.allclose(p.grad.norm(), torch.tensor(2.0 * (rank + 1)) / n2, rtol=1e-4)
    q = nn.Parameter(torch.zeros(2))
    assert float(clip_grad_norm_([q], 1.0)) == 0.0            # no gradients at all
    p.grad = torch.full((4,), float("nan") if rank == 3 else 1.0)
    with pytest.raises(RuntimeError):
        clip_grad_norm_([p], 1.0, error_if_nonfinite=True)    # every rank sees the agreed norm

    from torchdistpackage_b200.parallel.tensor_parallel import tp_utils
    tp_utils.set_tp_group(tdp.tpc.get_group("tensor"))
    tp = tdp.tpc.get_tp_rank()
    col = ColParallelLinear(4, 12, bias=True)
    full_w = torch.arange(48.0).view(4, 12)                   # [fin, 3*dim]: q | k | v thirds
    full_b = torch.arange(12.0)
    with torch.no_grad():
        col.init_weight_from_full_attn(full_w, full_b)
    want = torch.cat([t.split(2, dim=-1)[tp] for t in full_w.split(4, dim=-1)], dim=-1)
    assert torch.equal(col.linear.weight.detach(), want)      # heads stay aligned per projection
    want_b = torch.cat([t.split(2)[tp] for t in full_b.split(4)])
    assert torch.equal(col.linear.bias.detach(), want_b)
    tp_utils.reset_tp_group()


def test_clip_grad_norm_variants_and_attention_weight_slicing():
    run_distributed(_w_clip_variants, 4)


# ------------------------------------------------------------------ fused SP autograd wiring (CPU twin)
def _w_tp_fused_twin(rank, world):
    """parallel/tensor_parallel/tp_fused.py: the autograd functions around the fused GEMM+collective
    kernels -- which tensors are saved, which product pairs with which collective in backward, the
    GELU derivative riding on the all-gather GEMM -- run here with the three kernel entry points
    replaced by their definitions in torch + gloo (all_gather -> matmul, matmul -> reduce_scatter,
    matmul -> all_reduce).  Outputs and ALL gradients must match plain autograd over the unfused
    formulation (what the reference computes: tp_utils.py:52-159, mlp.py:69-78)."""
    import torch.nn.functional as F
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.ops import linear as L
    from torchdistpackage_b200.parallel.tensor_parallel import tp_fused as tf

    def ag(x):                                              # all-gather along dim 0
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous())
        return torch.cat(parts, 0)

    def rs(y):                                              # reduce-scatter along dim 0 (sum)
        y = y.clone()
        dist.all_reduce(y)
        return y.chunk(world, 0)[rank].contiguous()

    def act_apply(z, code):
        if code == L.ACT_NONE:
            return z
        return F.gelu(z, approximate="tanh" if code == L.ACT_GELU_TANH else "none")

    def dact(z, code):
        zf = z.detach().clone().requires_grad_(True)
        base = L.ACT_GELU_TANH if code == L.ACT_DGELU_TANH else L.ACT_GELU_ERF
        with torch.enable_grad():
            a = act_apply(zf, base)
        return torch.autograd.grad(a, zf, torch.ones_like(a))[0]

    class Token:                                            # what _gathered_or_regather inspects
        pass

    def fake_ag_gemm(ctx, name, x_shard, w, trans_b, bias=None, act=0, aux_in=None, want_aux_out=False):
        full = ag(x_shard)
        z = full @ (w.t() if trans_b else w)
        if bias is not None:
            z = z + bias
        if act in (L.ACT_DGELU_TANH, L.ACT_DGELU_ERF):
            out = z * dact(aux_in, act)
        else:
            out = act_apply(z, act)
        full._tdp_token = Token()
        return out, full, (z if want_aux_out else None)

    def fake_gemm_rs(ctx, name, a, w, trans_b, bias=None, residual=None):
        y = rs(a @ (w.t() if trans_b else w))
        if bias is not None:
            y = y + bias
        return y if residual is None else y + residual

    def fake_gemm_ar(ctx, name, a, w, trans_b, bias=None):
        y = a @ (w.t() if trans_b else w)
        dist.all_reduce(y)
        return y if bias is None else y + bias

    saved = (tf._ag_gemm, tf._gemm_rs, tf._gemm_ar, tf._gathered_or_regather)
    saved_l = (L.gemm, L.colsum)
    tf._ag_gemm, tf._gemm_rs, tf._gemm_ar = fake_ag_gemm, fake_gemm_rs, fake_gemm_ar
    tf._gathered_or_regather = lambda fctx, gathered, token, shard: gathered
    # the plain GEMM / column-sum entry points of the native extension, by definition
    L.gemm = lambda a, b, trans_a=False, trans_b=False, **kw: \
        (a.t() if trans_a else a) @ (b.t() if trans_b else b)
    L.colsum = lambda x, out_dtype=None: x.sum(0)
    try:
        torch.manual_seed(5)                               # same weights everywhere ...
        T, K, Hh = 8, 6, 10
        w1, b1 = torch.randn(K, Hh), torch.randn(Hh)
        w2, b2 = torch.randn(Hh, K), torch.randn(K)
        # ... tensor-parallel shards of them on this rank
        w1s, b1s = w1.chunk(world, 1)[rank].contiguous(), b1.chunk(world)[rank].contiguous()
        w2s = w2.chunk(world, 0)[rank].contiguous()
        torch.manual_seed(100 + rank)
        x_shard = torch.randn(T // world, K)                # sequence shard of the activations
        g_shard = torch.randn(T // world, K)                # incoming gradient (shard)

        def leaves(*ts):
            return [t.detach().clone().requires_grad_(True) for t in ts]

        def grads_of(y, g, params):
            return torch.autograd.grad(y, params, g, allow_unused=True)

        # (1) sp_mlp  ==  RS(gelu(AG(x) @ W1 + b1) @ W2) + b2
        pa = leaves(x_shard, w1s, b1s, w2s, b2)
        ya = tf.sp_mlp(None, pa[0], pa[1], pa[2], pa[3], pa[4], act="gelu_tanh")
        pb = leaves(x_shard, w1s, b1s, w2s, b2)

        class AG(torch.autograd.Function):                  # reference formulation with its own
            @staticmethod                                   # collectives' autograd rules
            def forward(ctx, t):
                return ag(t)

            @staticmethod
            def backward(ctx, gfull):
                return rs(gfull)

        class RS(torch.autograd.Function):
            @staticmethod
            def forward(ctx, t):
                return rs(t)

            @staticmethod
            def backward(ctx, gsh):
                return ag(gsh)

        yb = RS.apply(F.gelu(AG.apply(pb[0]) @ pb[1] + pb[2], approximate="tanh") @ pb[3]) + pb[4]
        assert torch.allclose(ya, yb, atol=1e-5)
        for name, ga, gb in zip(("dx", "dw1", "db1", "dw2", "db2"), grads_of(ya, g_shard, pa),
                                grads_of(yb, g_shard, pb)):
            assert torch.allclose(ga, gb, atol=1e-4), name

        # (2) ag_linear / linear_rs as separate ops
        pa, pb = leaves(x_shard, w1s, b1s), leaves(x_shard, w1s, b1s)
        ya = tf.ag_linear(None, pa[0], pa[1], pa[2])
        yb = AG.apply(pb[0]) @ pb[1] + pb[2]
        gfull = torch.randn_like(yb)
        assert torch.allclose(ya, yb, atol=1e-5)
        for name, ga, gb in zip(("dx", "dw", "db"), grads_of(ya, gfull, pa), grads_of(yb, gfull, pb)):
            assert torch.allclose(ga, gb, atol=1e-4), name
        a_full = torch.randn(T, Hh // world)
        pa, pb = leaves(a_full, w2s, b2), leaves(a_full, w2s, b2)
        ya = tf.linear_rs(None, pa[0], pa[1], pa[2])
        yb = RS.apply(pb[0] @ pb[1]) + pb[2]
        assert torch.allclose(ya, yb, atol=1e-5)
        for name, ga, gb in zip(("da", "dw", "db"), grads_of(ya, g_shard, pa), grads_of(yb, g_shard, pb)):
            assert torch.allclose(ga, gb, atol=1e-4), name

        # (3) linear_ar (row-parallel without SP): y = all_reduce(a @ W) + b, identity in backward
        pa, pb = leaves(a_full, w2s, b2), leaves(a_full, w2s, b2)
        ya = tf.linear_ar(None, pa[0], pa[1], pa[2])

        class AR(torch.autograd.Function):
            @staticmethod
            def forward(ctx, t):
                t = t.clone()
                dist.all_reduce(t)
                return t

            @staticmethod
            def backward(ctx, g_):
                return g_
        yb = AR.apply(pb[0] @ pb[1]) + pb[2]
        gy = torch.randn(T, K)
        assert torch.allclose(ya, yb, atol=1e-5)
        for name, ga, gb in zip(("da", "dw", "db"), grads_of(ya, gy, pa), grads_of(yb, gy, pb)):
            assert torch.allclose(ga, gb, atol=1e-4), name
    finally:
        tf._ag_gemm, tf._gemm_rs, tf._gemm_ar, tf._gathered_or_regather = saved
        L.gemm, L.colsum = saved_l
    del tdp


def test_fused_sequence_parallel_autograd_wiring_cpu_twin():
    run_distributed(_w_tp_fused_twin, 2)


# ------------------------------------------------------------------ differential: TP block vs the reference on gloo
def _w_tp_block_vs_reference(rank, world):
    """The unmodified reference's ParallelBlock (non-SP: its collectives are plain all_reduce, which
    gloo has) and this package's, both initialised from the same serial block with
    ``init_from_full``: same forward, same weight gradients.  (Input gradients are NOT compared:
    the reference's column-parallel layers do not all-reduce them, SURVEY.md 2.6 #11 -- ours must
    match the serial block instead, which test_tp_block_matches_serial checks.)"""
    import os
    import sys
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    sys.path.insert(0, ref_dir)
    try:
        from torchdistpackage.parallel.tensor_parallel.transformer import Block as RBlock, \
            ParallelBlock as RParallelBlock
    finally:
        sys.path.remove(ref_dir)
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.parallel import Block, ParallelBlock
    tdp.fix_rand(0)
    dim, heads = 32, 4
    serial_ref = RBlock(dim, num_heads=heads)
    with torch.no_grad():
        for p in serial_ref.parameters():              # torch.rand weights blow activations up
            p.mul_(0.1)
    serial = Block(dim, num_heads=heads)
    serial.load_state_dict(serial_ref.state_dict())    # same names and shapes
    theirs = RParallelBlock(dim, num_heads=heads, sequence_parallel=False)
    ours = ParallelBlock(dim, num_heads=heads, sequence_parallel=False)
    theirs.init_from_full(serial_ref)
    ours.init_from_full(serial)
    assert [(n, tuple(p.shape)) for n, p in theirs.named_parameters()] == \
           [(n, tuple(p.shape)) for n, p in ours.named_parameters()]
    for (n, p), (_, q) in zip(theirs.named_parameters(), ours.named_parameters()):
        if n.endswith("bias") and not n.startswith("ln"):
            continue        # the reference leaves row-parallel / qkv biases at zero (init_from_full
                            # copies weights only); ours copies them from the serial block
        assert torch.equal(p, q), n
    with torch.no_grad():   # align the biases: zero everywhere, as in the reference after init
        for blk in (theirs, ours):
            for n, p in blk.named_parameters():
                if n.endswith("bias") and not n.startswith("ln"):
                    p.zero_()
    torch.manual_seed(3)
    x = torch.randn(2, 6, dim)
    y_t = theirs(x.clone())
    y_o = ours(x.clone())
    scale = float(y_t.detach().abs().max())
    assert float((y_t - y_o).detach().abs().max()) <= 2e-5 * scale + 1e-6
    g = torch.randn_like(y_t)
    y_t.backward(g)
    y_o.backward(g)
    checked = 0
    for (n, p), (_, q) in zip(theirs.named_parameters(), ours.named_parameters()):
        # Only the MLP weights see the same upstream gradient in both implementations: everything
        # before the MLP receives, in the reference, the *partial* input gradient of the
        # column-parallel fc1 (no all-reduce in its backward, defect #11), so its attention / LN
        # gradients are not what the serial block gives -- ours are (test_tp_block_matches_serial).
        if not (n.startswith("mlp.") and n.endswith("weight")):
            continue
        s_ = float(p.grad.abs().max())
        assert float((p.grad - q.grad).abs().max()) <= 5e-5 * s_ + 1e-6, n
        checked += 1
    assert checked == 2
    # and the reference's attention-side gradients really are different from the serial truth
    xs = x.clone().requires_grad_(True)
    ys = serial(xs)
    ys.backward(g)
    tp = rank
    full = serial.attn.proj.weight.grad                      # [dim, dim]; row-parallel shard = rows
    mine = full[tp * (dim // world):(tp + 1) * (dim // world)]
    ours_g = dict(ours.named_parameters())["attn.proj.linear.weight"].grad
    assert float((ours_g - mine).abs().max()) <= 5e-5 * float(mine.abs().max()) + 1e-6


def test_parallel_block_matches_the_reference_on_gloo():
    import os
    if not os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                      "baseline", "_ref", "torchdistpackage")):
        pytest.skip("reference arm not installed (baseline/_ref)")
    run_distributed(_w_tp_block_vs_reference, 2)
