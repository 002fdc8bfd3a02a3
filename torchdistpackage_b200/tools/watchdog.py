"""Host-side hang / failure detection for long training jobs.

The reference has no in-job failure detection (only the 100 s ``new_group`` timeout
and the SLURM resubmitter).  On one NVSwitch node the failure that matters is a *hang*: one rank
stops launching work (data loader stall, host exception swallowed by a thread, a peer died) and
every other rank spins in a collective.  Device-side spins in this package already carry a 10 s
watchdog that traps with the wait site; this module is the host half:

``StepWatchdog(timeout_s)`` runs a daemon thread; the training loop calls ``tick(step)`` once per
step.  If no tick arrives for ``timeout_s`` the watchdog dumps the Python stack of every thread
(``faulthandler``), optionally the per-rank last-seen step of all ranks (a tiny TCPStore the
ranks write into, so the slow rank is named), calls ``on_hang`` and, if ``abort=True``, kills the
process with a distinctive exit code so ``tools/slurm_job_monitor.py`` (or torchrun's restart
policy) can restart from the last checkpoint.
"""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Callable, Dict, Optional

HANG_EXIT_CODE = 87


class StepWatchdog:
    def __init__(self, timeout_s: float = 600.0, abort: bool = False,
                 on_hang: Optional[Callable[[dict], None]] = None, store=None,
                 rank: Optional[int] = None, world: Optional[int] = None,
                 poll_s: Optional[float] = None, stream=None):
        self.timeout_s = float(timeout_s)
        self.abort = bool(abort)
        self.on_hang = on_hang
        self.store = store                      # optional torch.distributed.Store (e.g. TCPStore)
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else int(world)
        self.poll_s = poll_s if poll_s is not None else min(5.0, max(0.05, self.timeout_s / 10))
        self.stream = stream or sys.stderr
        self._last_tick = time.monotonic()
        self._last_step = -1
        self._stop = threading.Event()
        self._fired = False
        self._thread: Optional[threading.Thread] = None

    # ------------------------------------------------------------------ training-loop side
    def start(self) -> "StepWatchdog":
        self._last_tick = time.monotonic()
        self._thread = threading.Thread(target=self._run, name="tdp-step-watchdog", daemon=True)
        self._thread.start()
        return self

    def tick(self, step: int) -> None:
        self._last_step = int(step)
        self._last_tick = time.monotonic()
        self._fired = False
        if self.store is not None:
            try:
                self.store.set(f"tdp_wd_step_{self.rank}", str(int(step)))
            except Exception:       # the store host may be the rank that died
                pass

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2 * self.poll_s + 1)

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()

    # ------------------------------------------------------------------ watchdog side
    def peer_steps(self) -> Dict[int, Optional[int]]:
        """Last step every rank reported (None: never reported / store unreachable)."""
        out: Dict[int, Optional[int]] = {}
        if self.store is None:
            return out
        for r in range(self.world):
            try:
                out[r] = int(self.store.get(f"tdp_wd_step_{r}").decode()) \
                    if self.store.check([f"tdp_wd_step_{r}"]) else None
            except Exception:
                out[r] = None
        return out

    def _report(self) -> dict:
        idle = time.monotonic() - self._last_tick
        info = {"rank": self.rank, "last_step": self._last_step, "idle_s": round(idle, 1),
                "peer_steps": self.peer_steps()}
        print(f"[tdp watchdog] rank {self.rank}: no training step for {idle:.1f} s "
              f"(last step {self._last_step}); peers: {info['peer_steps']}", file=self.stream, flush=True)
        self._dump_stacks()
        return info

    def _dump_stacks(self) -> None:
        """Python stack of every thread (where is the training loop stuck?)."""
        try:
            faulthandler.dump_traceback(file=self.stream, all_threads=True)
            return
        except Exception:           # stream without a file descriptor: format the frames ourselves
            pass
        import traceback
        names = {t.ident: t.name for t in threading.enumerate()}
        for ident, frame in sys._current_frames().items():
            print(f"Thread {names.get(ident, '?')} ({ident}):", file=self.stream)
            print("".join(traceback.format_stack(frame)), file=self.stream, flush=True)

    def _run(self) -> None:
        while not self._stop.wait(self.poll_s):
            if self._fired or time.monotonic() - self._last_tick < self.timeout_s:
                continue
            self._fired = True
            info = self._report()
            if self.on_hang is not None:
                try:
                    self.on_hang(info)
                except Exception as e:      # never let the callback kill the watchdog
                    print(f"[tdp watchdog] on_hang raised {e!r}", file=self.stream, flush=True)
            if self.abort:
                os._exit(HANG_EXIT_CODE)
