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

    return counted_step, launches, model.cfg


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
        emit_json(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
