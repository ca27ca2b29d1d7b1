"""Sample GPU clocks / throttle reasons with nvidia-smi while a timed region runs."""
from __future__ import annotations

import statistics
import subprocess
import threading
import time

_QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
          "clocks_event_reasons.sw_power_cap")
_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 50):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.rows = []
        self._proc = None
        self._thread = None

    def start(self):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + _QUERY,
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self._proc = None
            return self
        self._thread = threading.Thread(target=self._read, daemon=True)
        self._thread.start()
        return self

    def _read(self):
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                self.rows.append((time.time(), float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
            except ValueError:
                continue

    def stop(self):
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        if self._thread is not None:
            self._thread.join(timeout=2)
        return self.summary()

    def summary(self, t0: float = None, t1: float = None):
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({name for r in rows for name, v in zip(_REASONS, r[4]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "power_w_max": max(r[3] for r in rows), "reasons": reasons, "samples": len(rows)}
