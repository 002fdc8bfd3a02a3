"""The scripts under examples/ are this package's counterpart of the reference's examples (which
are its only tests): each one must run to its "OK" line on CPU with gloo,
launched the way a user launches it (torchrun, one process per rank)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (script, ranks, text the output must contain)
DISTRIBUTED = [
    ("test_ddp.py", 2, "NaiveDDP == torch DDP"),
    ("test_zero_optim.py", 2, "ZeRO == DDP+Adam"),
    ("test_shard_ema.py", 2, "ShardedEMA == full EMA"),
    ("hybrid_zero.py", 4, "step 2 loss"),
    ("model_parallel/test_attn.py", 2, "OK"),
    ("model_parallel/test_tpmlp.py", 2, "sequence_parallel=True"),
    ("model_parallel/test_transformer.py", 2, "OK"),
    ("model_parallel/test_pipeline.py", 4, "pipeline example done"),
    ("moe/train_moe.py", 4, "step 4 loss"),
]
SINGLE = [
    ("understand_ops/norm_from_scratch.py", "LayerNorm"),
    ("tile_attention.py", "causal=True: tiled attention fwd/bwd match SDPA"),
    ("fx_profile_split.py", "cut after node"),
    ("profile/test_profile.py", "level: 1"),
]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _env():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    for k in list(env):
        if k.startswith(("SLURM_", "TORCHELASTIC")) or k in ("RANK", "WORLD_SIZE", "LOCAL_RANK",
                                                             "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k)
    return env


@pytest.fixture(scope="module")
def launched():
    """All example scripts are started together (each is its own torchrun job on its own port) and
    collected once: the wall time of this file is the slowest example, not their sum."""
    jobs = []
    for script, ranks, _ in DISTRIBUTED:
        jobs.append((script, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                              f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
                              "--master-port", str(_free_port()),
                              os.path.join(ROOT, "examples", script), "--cpu"]))
    for script, _ in SINGLE:
        jobs.append((script, [sys.executable, os.path.join(ROOT, "examples", script)]))
    out = {}
    width = 5                                   # jobs in flight (each is 1-4 processes)
    for i in range(0, len(jobs), width):
        procs = {script: subprocess.Popen(cmd, cwd=ROOT, env=_env(), stdout=subprocess.PIPE,
                                          stderr=subprocess.PIPE, text=True)
                 for script, cmd in jobs[i:i + width]}
        for script, p in procs.items():
            try:
                so, se = p.communicate(timeout=600)
                out[script] = (p.returncode, so, se)
            except subprocess.TimeoutExpired:
                p.kill()
                so, se = p.communicate()
                out[script] = (-9, so, se + "\n[timeout]")
    return out


@pytest.mark.parametrize("script,ranks,expect", DISTRIBUTED, ids=[d[0] for d in DISTRIBUTED])
def test_distributed_example_runs_on_gloo(launched, script, ranks, expect):
    rc, so, se = launched[script]
    assert rc == 0, (so[-1500:], se[-2500:])
    assert expect in so, so[-1500:]


@pytest.mark.parametrize("script,expect", SINGLE, ids=[d[0] for d in SINGLE])
def test_single_process_example_runs_on_cpu(launched, script, expect):
    rc, so, se = launched[script]
    assert rc == 0, (so[-1500:], se[-2500:])
    assert expect in so, so[-1500:]


def test_baseline_config_1_cpu_gloo_bench_line():
    """BASELINE config #1 (NaiveDDP on the reference's 2-layer MLP, world 2, CPU / gloo) through
    ``bench.py --config cpu``: one JSON line with a throughput and an exact gradient check."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "bench.py"), "--config", "cpu", "--steps", "50", "--warmup", "5"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-500:]
    rec = json.loads(lines[0])
    assert rec["impl"] == "ours" and rec["n_procs"] == 2 and rec["value"] > 0
    assert rec["grad_check_max_abs"] < 1e-6
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "cpu", "--impl", "reference"],
                       cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "unavailable" in json.loads(r.stdout.strip().splitlines()[-1])
