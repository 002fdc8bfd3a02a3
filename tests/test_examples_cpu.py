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


@pytest.mark.parametrize("script,ranks,expect", DISTRIBUTED, ids=[d[0] for d in DISTRIBUTED])
def test_distributed_example_runs_on_gloo(script, ranks, expect):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "examples", script), "--cpu"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2500:])
    assert expect in r.stdout, r.stdout[-1500:]


@pytest.mark.parametrize("script,expect", SINGLE, ids=[d[0] for d in SINGLE])
def test_single_process_example_runs_on_cpu(script, expect):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], cwd=ROOT, env=_env(),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2500:])
    assert expect in r.stdout, r.stdout[-1500:]
