"""Submit a SLURM batch script and keep it alive: poll ``sacct`` and re-``sbatch`` when the job
leaves the healthy states (reference: tools/slurm_job_monitor.py:1-132, tools/sbatch.sh).

    python -m torchdistpackage_b200.tools.slurm_job_monitor --cfg job.sh [--interval 10] [--max-restarts 20]

Resumption from a checkpoint is the job script's responsibility (see dist/model_parallel_ckpt.py
and ``Bf16ZeroOptimizer.state_dict``).
"""
from __future__ import annotations

import argparse
import re
import subprocess
import time
from typing import Callable, Optional

HEALTHY = {"RUNNING", "PENDING", "COMPLETED", "COMPLETING", "CONFIGURING"}


def _run(cmd) -> str:
    return subprocess.run(cmd, capture_output=True, text=True, check=False).stdout


def submit(script: str, runner: Callable = _run) -> Optional[str]:
    out = runner(["sbatch", script])
    m = re.search(r"Submitted batch job (\d+)", out)
    return m.group(1) if m else None


def job_state(job_id: str, runner: Callable = _run) -> str:
    out = runner(["sacct", "-j", job_id, "--format=State", "--noheader", "-X"])
    toks = out.split()
    return toks[0].rstrip("+") if toks else "UNKNOWN"


def monitor_job(script: str, interval: float = 10.0, max_restarts: int = 100,
                runner: Callable = _run, sleep: Callable = time.sleep) -> int:
    """Returns the number of (re)submissions made."""
    submissions = 0
    job_id = submit(script, runner)
    submissions += 1
    print(f"[monitor] submitted {script} as job {job_id}", flush=True)
    while job_id is not None:
        sleep(interval)
        state = job_state(job_id, runner)
        if state == "COMPLETED":
            print(f"[monitor] job {job_id} completed", flush=True)
            break
        if state not in HEALTHY and state != "UNKNOWN":
            if submissions > max_restarts:
                print("[monitor] restart budget exhausted", flush=True)
                break
            print(f"[monitor] job {job_id} is {state}: resubmitting", flush=True)
            job_id = submit(script, runner)
            submissions += 1
    return submissions


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True, help="sbatch script to run and watch")
    ap.add_argument("--interval", type=float, default=10.0)
    ap.add_argument("--max-restarts", type=int, default=100)
    a = ap.parse_args()
    monitor_job(a.cfg, a.interval, a.max_restarts)


if __name__ == "__main__":
    main()
