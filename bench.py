#!/usr/bin/env python
"""Flagship benchmark: GPT-2 small, pure data parallel (NaiveDdp), bf16, synthetic tokens.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3
    python bench.py --impl reference ...      # the unmodified reference (baseline/_ref) arm

Metric (BASELINE.json): transformer tokens/sec, whole job, device-timed (CUDA events), max over
ranks; weak scaling (fixed per-GPU micro-batch).  One JSON line on rank 0.

What a step is (nothing skipped): zero grads -> forward (loss) -> backward with bucketed gradient
all-reduce overlapped on a side stream -> reduce_gradients() -> AdamW step.
  * ours:       torchdistpackage_b200 GPT-2 (tcgen05 GEMMs with fused epilogues, fused LN / CE),
                NaiveDDP over NVLS symmetric-memory buckets, BucketAdamW (one fused launch per bucket).
  * reference:  plain-torch GPT-2 of the same architecture wrapped in the reference's NaiveDDP
                (baseline/ref_bench.py) + torch.optim.AdamW(fused=True).
``value`` is measured with the batch already resident on the device; ``e2e`` repeats the same K
steps through the public API with a per-step pinned-host -> device copy of the tokens and a
device -> host read of the loss.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MICRO_BATCH = 16      # sequences per GPU per step
SEQ_LEN = 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="small", choices=["small", "medium", "tiny"])
    ap.add_argument("--micro-batch", type=int, default=MICRO_BATCH)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="ours: run the step eagerly")
    ap.add_argument("--other-configs", default="auto", choices=["auto", "on", "off"],
                    help="after the headline run, also time BASELINE configs #3/#4/#5 (both arms) "
                         "outside the timed region; auto = only at --gpus 8")
    ap.add_argument("--no-checks", action="store_true",
                    help="skip the untimed gradient check / exposed-communication measurement")
    ap.add_argument("--config", default="dp", choices=["dp", "tp", "moe", "mixed", "cpu"],
                    help="dp = the headline (BASELINE config #2, this file); tp / moe / mixed = "
                         "BASELINE configs #3 / #4 / #5, handed to scripts/bench_{tp,moe,mixed}.py "
                         "with the same --impl / --steps / --warmup (one JSON line each); cpu = "
                         "config #1, the CPU / gloo plumbing check (scripts/bench_cpu_mlp.py)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        # NVML in-process (20 ms period) when available, else an `nvidia-smi -lms` child
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:        # CUDA_VISIBLE_DEVICES-proof: look the device up by UUID
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = pynvml
            self.samples = []
            self.stop_flag = False
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def _stop_nvml(self) -> dict:
        self.stop_flag = True
        self.t.join(timeout=2)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        reasons = set()
        for _, _, _, rs in self.samples:
            for bit, name in names.items():
                if rs & bit:
                    reasons.add(name)
        sm = sorted(x[0] for x in self.samples)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None,
                "sm_max_mhz": float(max(x[1] for x in self.samples)) if sm else None,
                "power_w_max": max(x[2] for x in self.samples) if sm else None,
                "samples": len(sm), "source": "nvml", "reasons": sorted(reasons)}

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if getattr(self, "nvml", None) is not None:
            return self._stop_nvml()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_ours(args, device, world):
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.models.gpt2 import build_gpt2
    from torchdistpackage_b200.ops.fused import BucketAdamW
    tdp.fix_rand(0, deterministic_cudnn=False)
    model = build_gpt2(args.model, device=device)
    pg = None
    if world > 1:
        tdp.tpc.verbose = False
        tdp.tpc.setup_process_groups([("data", world)])
        pg = tdp.tpc.get_group("data")
    ddp = tdp.NaiveDDP(model, sync=False, gradient_as_bucket_view=True, bucket_cap_mb=25,
                       process_group=pg)
    opt = BucketAdamW(ddp, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    native = tdp.ops.native()

    def eager_step(tokens, targets):
        opt.zero_grad()
        loss = ddp(tokens, targets)
        loss.backward()
        ddp.reduce_gradients()
        opt.step()
        return loss

    state = {"graph": None, "launches_per_replay": 0}

    def step(tokens, targets):
        """Public training step: the eager step captured once into a CUDA graph, then replayed."""
        if args.no_graph:
            return eager_step(tokens, targets)
        if state["graph"] is None:
            from torchdistpackage_b200.ops.graph import GraphedStep
            before = int(native.launch_count())
            state["graph"] = GraphedStep(eager_step, (tokens, targets), warmup=2)
            # kernels of ours inside one captured step (warm-up iterations + capture = 3 steps)
            state["launches_per_replay"] = (int(native.launch_count()) - before) // 3
        return state["graph"](tokens, targets)

    replays = {"n": 0}

    def counted_step(tokens, targets):
        replays["n"] += 1
        return step(tokens, targets)

    def launches():
        if args.no_graph:
            return int(native.launch_count()) if native is not None else 0
        # a replay re-launches every captured kernel; the host-side counter does not see replays
        return replays["n"] * state["launches_per_replay"]

    build_ours.ctx = dict(ddp=ddp, opt=opt, eager_step=eager_step, args=args)
    return counted_step, launches, model.cfg


def grad_check_ours(step, batch, world, device):
    """Untimed proof that the gradient the timed steps used is the data-parallel average.

    One more step of the *timed* path (CUDA-graph replay, direct weight-gradient route into the
    symmetric buckets, NVLS all-reduce kernels) on ``batch`` with lr = 0 (weights frozen); the
    reduced buckets are kept.  Then the same batch runs through the eager step with the bucket
    reduction disabled, which leaves every rank's *local* gradient in the buckets; those are
    averaged with plain ``dist.all_reduce`` (NCCL) on a copy and compared."""
    import torch.distributed as dist
    ctx = build_ours.ctx
    ddp, opt = ctx["ddp"], ctx["opt"]
    red = ddp.reducer
    lr0 = float(opt.param_groups[0]["lr"])
    opt.set_lr(0.0)
    tokens, targets = batch[:, :-1], batch[:, 1:]
    fused = bool(getattr(opt, "fused_comm", False))
    if os.environ.get("TDP_BENCH_GPU_LAG"):
        # debugging aid: let the host run far ahead of the device (as it does on a loaded 8-GPU
        # box), which exposes any missing device-side ordering between the streams
        torch.cuda._sleep(int(float(os.environ["TDP_BENCH_GPU_LAG"]) * 1.9e9))
    if fused:
        # fused reduce-scatter -> AdamW -> all-gather kernels never materialise the averaged
        # gradient; the same kernels run once (eagerly) with their write-back flag on
        opt.write_back_grad = True
        try:
            ctx["eager_step"](tokens.contiguous(), targets.contiguous())
        finally:
            opt.write_back_grad = False
    else:
        step(tokens, targets)
    torch.cuda.synchronize()
    got = [b.payload().float().clone() for b in red.buckets]
    orig = red._reduce_bucket
    red._reduce_bucket = lambda bucket: setattr(bucket, "reduced", True)
    try:
        ctx["eager_step"](tokens.contiguous(), targets.contiguous())
        torch.cuda.synchronize()
    finally:
        red._reduce_bucket = orig
    worst_max, worst_l2, sym = 0.0, 0.0, 0
    detail = []
    for b, g in zip(red.buckets, got):
        local = b.payload().float().clone()
        ref = local.clone()
        dist.all_reduce(ref, group=b.group)
        ref /= world
        e_max = float((g - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
        e_l2 = float((g - ref).norm() / ref.norm().clamp_min(1e-20))
        worst_max, worst_l2 = max(worst_max, e_max), max(worst_l2, e_l2)
        sym += int(b.symm is not None)
        if os.environ.get("TDP_BENCH_GRAD_DETAIL"):
            bad = (g - ref).abs() > 0.05 * ref.abs().max()
            nb = int(bad.sum())
            idx = bad.nonzero().flatten()
            detail.append(dict(bucket=b.index, numel=int(g.numel()), names=b.names[:2], rel=e_max, l2=e_l2,
                               n_bad=nb, first_bad=int(idx[0]) if nb else -1, last_bad=int(idx[-1]) if nb else -1,
                               got_vs_local_l2=float((g - local).norm() / local.norm().clamp_min(1e-20)),
                               got_norm=float(g.norm()), ref_norm=float(ref.norm()), local_norm=float(local.norm())))
    if detail:
        import torch.distributed as _d
        print(f"[grad detail rank {_d.get_rank()}] " + json.dumps(detail), file=sys.stderr, flush=True)
    opt.set_lr(lr0)
    # replicas must hold bit-identical parameters after all those steps (the fused path
    # multicasts them; a lost store would show up here)
    chk = torch.stack([st["flat_p"].view(torch.int16).to(torch.int64).sum() for st in opt.state])
    hi_, lo_ = chk.clone(), chk.clone()
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    params_identical = bool((hi_ == lo_).all().item())
    t = torch.tensor([worst_max, worst_l2], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"grad_check_rel": float(t[0]), "grad_check_rel_l2": float(t[1]),
            "params_identical_across_ranks": params_identical, "fused_reduce_optimizer": fused,
            "grad_check_buckets": len(red.buckets), "grad_check_symmetric_buckets": sym,
            "grad_check_how": ("the step's own reduction kernels (direct wgrad -> NVLS reduce; fused "
                               "optimizer mode: run eagerly with the averaged gradient written "
                               "back) vs local grads of a second backward pass averaged by NCCL "
                               "all_reduce, same batch, lr=0; max over buckets and ranks of "
                               "max|diff|/max|ref| and of the L2 ratio (the two passes differ by "
                               "their own atomics: bf16 row scatter of the tied embedding, fp32 "
                               "column sums)")}


def exposed_comm_ours(dev_batches, K, barrier, max_over_ranks, ms_with_comm):
    """BASELINE metric, second half: communication left exposed after overlap = step time with
    the bucket all-reduces minus the time of the identical step with the collective skipped
    (a second CUDA graph captured with the reduction disabled; everything else identical)."""
    from torchdistpackage_b200.ops.graph import GraphedStep
    ctx = build_ours.ctx
    red = ctx["ddp"].reducer
    opt = ctx["opt"]
    orig = red._reduce_bucket
    red._reduce_bucket = lambda bucket: setattr(bucket, "reduced", True)
    # the replicas see un-averaged gradients in this run: freeze the weights (lr = 0 -- the
    # optimizer kernels still run, with identical cost) so that they cannot drift apart
    lr0 = float(opt.param_groups[0]["lr"])
    opt.set_lr(0.0)
    try:
        b0 = dev_batches[0]
        if ctx["args"].no_graph:
            g = ctx["eager_step"]
        else:
            g = GraphedStep(ctx["eager_step"], (b0[:, :-1], b0[:, 1:]), warmup=2)
        for i in range(3):
            b = dev_batches[i % len(dev_batches)]
            g(b[:, :-1], b[:, 1:])
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(K):
            b = dev_batches[i % len(dev_batches)]
            g(b[:, :-1], b[:, 1:])
        e.record()
        barrier()
        ms = max_over_ranks(s.elapsed_time(e)) / K
    finally:
        red._reduce_bucket = orig
        opt.set_lr(lr0)
    del g
    return {"exposed_comm_ms": ms_with_comm - ms, "ms_per_step_without_collective": ms}


OTHER_CONFIGS = [
    ("tp", "scripts/bench_tp.py", "#3 4-layer transformer h=4096 TP=N + SP"),
    ("moe", "scripts/bench_moe.py", "#4 MoE 8-expert transformer EP x moe-DP"),
    ("mixed", "scripts/bench_mixed.py", "#5 GPT-2 medium DP x PP=2 x TP=2 + SP, ZeRO, 1F1B"),
]


def _agree_to_run(rank: int, tag: str, go: bool, wait_s: float = 90.0) -> bool:
    """All ranks of this node take rank 0's decision (the process group is already gone, and a
    rank that starts a side run alone would sit in its rendezvous until the timeout): rank 0
    publishes it as a file, the others wait for the file."""
    path = os.path.join("/tmp", f"tdp_bench_{tag}")
    if rank == 0:
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            f.write("go" if go else "skip")
        os.replace(tmp, path)
        return go
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < wait_s:
        try:
            with open(path) as f:
                return f.read().strip() == "go"
        except OSError:
            time.sleep(0.05)
    return False


def run_other_configs(rank, world):
    """BASELINE configs #3 / #4 / #5, both arms, each as a fresh set of ``world`` processes (this
    rank spawns its counterpart with the same RANK / LOCAL_RANK on a new rendezvous port), timed by
    the scripts themselves (CUDA events, max over ranks) -- outside the headline timed region.
    The whole section has a wall-clock budget (``TDP_BENCH_OTHER_BUDGET_S``, default 420 s; a
    healthy side run takes about half a minute): once it is spent the remaining runs are
    skipped and reported as such, so a stuck side config cannot hold the headline line back
    for long.  Returns {name: {ours_ms, reference_ms, ratio, ...}} on rank 0."""
    base_port = int(os.environ.get("MASTER_PORT", "29500"))
    budget = float(os.environ.get("TDP_BENCH_OTHER_BUDGET_S", "420"))
    t_begin = time.perf_counter()
    tag = f"{base_port}_{os.getppid()}"      # the launcher's pid: same on every rank, new per launch
    out = {}
    idx = 0
    for name, script, title in OTHER_CONFIGS:
        if name == "mixed" and world % 4 != 0:
            continue
        rec = {"config": title}
        for impl in ("reference", "ours"):
            idx += 1
            left = budget - (time.perf_counter() - t_begin)
            if not _agree_to_run(rank, f"{tag}_{idx}", left > 30.0):
                if rank == 0:
                    rec[impl + "_error"] = "skipped: time budget of the side configs spent"
                continue
            env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC")}
            env["MASTER_PORT"] = str(base_port + 100 + idx)
            env["MASTER_ADDR"] = "127.0.0.1"
            res = None
            try:
                cp = subprocess.run([sys.executable, os.path.join(ROOT, script), "--impl", impl],
                                    env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                    text=True, timeout=180)
                for ln in cp.stdout.splitlines():
                    if ln.startswith("{"):
                        res = json.loads(ln)
                if res is None and rank == 0:
                    rec[impl + "_error"] = (cp.stderr or "")[-300:]
            except Exception as ex:       # a failed side config must not take the headline down
                if rank == 0:
                    rec[impl + "_error"] = f"{type(ex).__name__}: {ex}"[:300]
            if res is not None:
                rec[impl + "_ms_per_step"] = res["ms_per_step"]
                rec[impl + "_tokens_per_s"] = res["tokens_per_s"]
                rec["shape"] = res.get("config", title) if impl == "ours" else rec.get("shape", res.get("config"))
        if "ours_ms_per_step" in rec and "reference_ms_per_step" in rec:
            rec["ratio"] = rec["reference_ms_per_step"] / rec["ours_ms_per_step"]
        out[name] = rec
    if rank == 0:
        out["wall_s"] = round(time.perf_counter() - t_begin, 1)
        for i in range(1, idx + 1):
            try:
                os.remove(os.path.join("/tmp", f"tdp_bench_{tag}_{i}"))
            except OSError:
                pass
    return out


_REAL_STDOUT_FD = None


def _quiet_stdout():
    """stdout must carry exactly ONE JSON line.  Native libraries (NCCL prints its version banner
    with printf) write to fd 1 directly, so fd 1 is pointed at stderr for the whole run and the
    JSON line is written to a saved duplicate of the real stdout."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT_FD if _REAL_STDOUT_FD is not None else 1, line)


def main():
    args = parse_args()
    if args.config != "dp":
        script = dict([(name, path) for name, path, _ in OTHER_CONFIGS] +
                      [("cpu", "scripts/bench_cpu_mlp.py")])[args.config]
        os.execv(sys.executable, [sys.executable, os.path.join(ROOT, script), "--impl", args.impl,
                                  "--steps", str(args.steps), "--warmup", str(args.warmup)])
    _quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        try:
            import ref_bench
        except Exception as e:  # reference not installed / importable on this box
            if rank == 0:
                emit_json({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"})
            return
        builder = ref_bench.build_reference
    else:
        builder = build_ours

    if not torch.cuda.is_available():
        if args.impl == "reference":
            if rank == 0:
                emit_json({"impl": "reference", "unavailable": "no CUDA device"})
            return
        raise SystemExit("bench.py needs a CUDA device (B200)")

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist
    import contextlib
    # stdout carries exactly ONE JSON line: library chatter (the reference prints from
    # setup_distributed / group creation) goes to stderr
    with contextlib.redirect_stdout(sys.stderr):
        if world > 1:
            if args.impl == "reference":
                ref_bench.init_distributed()
            else:
                import torchdistpackage_b200 as tdp
                tdp.setup_distributed("nccl")
        assert world == args.gpus or world == 1 and args.gpus == 1, \
            f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
        step, launches, cfg = builder(args, device, world)
    B, S = args.micro_batch, cfg.seq_len
    K, W = args.steps, args.warmup
    gen = torch.Generator().manual_seed(1234 + rank)
    # K+W distinct synthetic batches in pinned host memory (tokens + next-token targets)
    n_batches = min(K + W, 8)
    host = [torch.randint(0, cfg.vocab_size, (B, S + 1), generator=gen).pin_memory()
            for _ in range(n_batches)]
    dev_batches = [h.to(device) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up (untimed)
    for i in range(W):
        b = dev_batches[i % n_batches]
        step(b[:, :-1], b[:, 1:])
    barrier()

    # ---------------- timed: K steps, inputs resident on the device
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launches()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(K):
        b = dev_batches[(W + i) % n_batches]
        loss = step(b[:, :-1], b[:, 1:])
    e.record()
    barrier()
    ms = max_over_ranks(s.elapsed_time(e))
    l1 = launches()
    final_loss = float(loss.item())

    # ---------------- timed: end to end (H2D of every step's tokens, D2H of every step's loss)
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for i in range(K):
            hb = host[(W + i) % n_batches]
            db = hb.to(device, non_blocking=True)           # pinned host -> device, this step
            loss = step(db[:, :-1], db[:, 1:])
            _ = loss.item()                                 # device -> host read of the result
        e2.record()
        barrier()
        ms2 = max_over_ranks(s2.elapsed_time(e2))
        wall2 = max_over_ranks((time.perf_counter() - t0) * 1e3)
        ms2 = max(ms2, 0.0)
        e2e = {"value": world * B * S * K / (ms2 / 1e3), "unit": "tokens/s",
               "ms_per_step": ms2 / K, "wall_ms_per_step": wall2 / K,
               "h2d_bytes_per_step": int(host[0].numel() * host[0].element_size()),
               "d2h_bytes_per_step": 4}
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- untimed extras (ours arm): exposed communication, gradient proof
    extras = {}
    if args.impl == "ours" and world > 1 and not args.no_checks:
        try:
            extras.update(exposed_comm_ours(dev_batches, K, barrier, max_over_ranks, ms / K))
        except Exception as ex:
            extras["exposed_comm_error"] = f"{type(ex).__name__}: {ex}"[:200]
        try:
            extras.update(grad_check_ours(step, dev_batches[0], world, device))
        except Exception as ex:
            extras["grad_check_error"] = f"{type(ex).__name__}: {ex}"[:200]
    do_other = args.other_configs == "on" or (args.other_configs == "auto" and world == 8)
    if do_other and world > 1 and args.impl == "ours":
        barrier()
        dist.destroy_process_group()
        build_ours.ctx = None
        torch.cuda.empty_cache()
        other = run_other_configs(rank, world)
        if rank == 0:
            extras["other_configs"] = other
        world_done = True
    else:
        world_done = False

    if rank == 0:
        tokens_per_s = world * B * S * K / (ms / 1e3)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_tflops = float(peaks.get("bf16_tflops_sustained", 1400.0))
        except Exception:
            peak_tflops = 1400.0
        mfu = tokens_per_s * cfg.flops_per_token() / world / (peak_tflops * 1e12)
        out = {
            "metric": "transformer tokens/sec (whole job, device-timed, max over ranks)",
            "value": tokens_per_s, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic tokens, random-init weights",
            "impl": args.impl,
            "config": {"model": f"gpt2-{args.model}", "global_batch": world * B, "seq_len": S,
                       "micro_batch_per_gpu": B, "parallelism": f"dp{world}",
                       "optimizer": "AdamW", "vocab": cfg.vocab_size,
                       "l2": "per-step working set (activations + weights, several GB) >> 126 MB "
                             "L2; distinct input batch every step; no explicit flush"},
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": (l1 - l0) if args.impl == "ours" else 0,
            "gpu_launches_per_step": ((l1 - l0) / K) if args.impl == "ours" else 0,
            "final_loss": final_loss,
            "model_flops_utilization_of_measured_cublas": mfu,
        }
        out.update(extras)
        emit_json(out)
    if world > 1 and not world_done:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
