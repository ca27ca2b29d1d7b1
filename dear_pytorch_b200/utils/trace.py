"""Chrome-trace (chrome://tracing / Perfetto) timeline of the DeAR state machine.

The reference ships a complete writer, dear/chrome_profiler.py:13-117 (queue + writer thread,
``B``/``E`` events per tensor/activity), that is never imported, and reads ``WFSGD_TIMELINE`` into
a field that is never used (dear/dopt_rb.py:198).  Here the timeline is wired in: set
``DEAR_TIMELINE=/path/trace.json`` (or pass a ``Timeline`` to ``attach``) and every bucket's
reduce-scatter / all-gather launch and every wait is recorded, one row per bucket.
"""
from __future__ import annotations

import json
import os
import queue
import threading
import time


class Timeline:
    def __init__(self, path: str, rank: int = 0):
        self.path = path
        self.rank = rank
        self._q: "queue.Queue" = queue.Queue()
        self._t0 = time.perf_counter()
        self._closed = False
        self._first = True
        self._f = open(path, "w")
        self._f.write("[\n")
        self._thread = threading.Thread(target=self._writer, daemon=True)
        self._thread.start()

    def _ts(self) -> float:
        return (time.perf_counter() - self._t0) * 1e6

    def _writer(self):
        while True:
            ev = self._q.get()
            if ev is None:
                break
            self._f.write(("" if self._first else ",\n") + json.dumps(ev))
            self._first = False
        self._f.write("\n]\n")
        self._f.close()

    def begin(self, row: str, activity: str, **args):
        self._q.put({"name": activity, "ph": "B", "ts": self._ts(), "pid": self.rank, "tid": row, "args": args})

    def end(self, row: str, activity: str, **args):
        self._q.put({"name": activity, "ph": "E", "ts": self._ts(), "pid": self.rank, "tid": row, "args": args})

    def instant(self, row: str, activity: str, **args):
        self._q.put({"name": activity, "ph": "i", "s": "t", "ts": self._ts(), "pid": self.rank, "tid": row, "args": args})

    def close(self):
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._thread.join(timeout=5)


def attach_backend(engine) -> None:
    """(Re-)instrument the engine's current backend (called again after every re-bucketing)."""
    timeline = getattr(engine, "timeline", None)
    if timeline is None:
        return
    be = engine.backend

    def wrap(name):
        fn = getattr(be, name)

        def inner(g, *a, **k):
            row = "bucket-%d" % g
            timeline.begin(row, name)
            try:
                return fn(g, *a, **k)
            finally:
                timeline.end(row, name)
        setattr(be, name, inner)
    for nm in ("reduce_scatter", "allgather_update", "wait_bucket"):
        wrap(nm)


def attach(engine, timeline: Timeline = None) -> Timeline:
    """Instrument a ``DearEngine`` (bucket launches, waits and ``step``)."""
    if timeline is None:
        path = os.environ.get("DEAR_TIMELINE")
        if not path:
            return None
        timeline = Timeline(path if engine.world == 1 else "%s.rank%d" % (path, engine.rank), engine.rank)
    engine.timeline = timeline
    attach_backend(engine)
    step = engine.step

    def timed_step():
        timeline.begin("optimizer", "step", num_steps=engine.num_steps)
        try:
            return step()
        finally:
            timeline.end("optimizer", "step")
    engine.step = timed_step
    return timeline
