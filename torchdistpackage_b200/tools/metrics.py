"""Minimal structured metrics: one JSON object per line, rank-0 only by default.

The reference logs with ``print`` only.  ``MetricsLogger`` keeps that zero-dependency
spirit but makes the numbers machine readable: ``log(step, loss=..., tokens_per_s=...)`` appends a
line ``{"step": .., "time": .., "rank": .., ...}`` to a ``.jsonl`` file (and optionally echoes it).
Tensors are converted with ``float(t)`` - pass detached scalars, ideally already on the host, so the
logger never adds a device sync of its own.
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Optional


class MetricsLogger:
    def __init__(self, path: Optional[str], rank: Optional[int] = None, all_ranks: bool = False,
                 echo: bool = False, flush_every: int = 1):
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        self.enabled = path is not None and (all_ranks or self.rank == 0)
        self.echo = echo
        self.flush_every = max(1, int(flush_every))
        self._n = 0
        self._f = None
        if self.enabled:
            if all_ranks:
                root, ext = os.path.splitext(path)
                path = f"{root}.rank{self.rank}{ext or '.jsonl'}"
            os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
            self._f = open(path, "a", buffering=1)
        self.path = path

    @staticmethod
    def _plain(v: Any):
        if hasattr(v, "item") and getattr(v, "numel", lambda: 2)() == 1:
            return float(v)
        if isinstance(v, (int, float, str, bool)) or v is None:
            return v
        if isinstance(v, (list, tuple)):
            return [MetricsLogger._plain(x) for x in v]
        if isinstance(v, dict):
            return {str(k): MetricsLogger._plain(x) for k, x in v.items()}
        return str(v)

    def log(self, step: int, **values) -> None:
        if not self.enabled:
            return
        rec = {"step": int(step), "time": round(time.time(), 3), "rank": self.rank}
        rec.update({k: self._plain(v) for k, v in values.items()})
        line = json.dumps(rec)
        self._f.write(line + "\n")
        self._n += 1
        if self._n % self.flush_every == 0:
            self._f.flush()
        if self.echo:
            print(line, flush=True)

    def close(self) -> None:
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
