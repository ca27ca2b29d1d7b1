"""Bayesian-optimisation tuner for the tensor-fusion buffer size (``dopt_rsag_bo`` behaviour).

Reference: dear/tuner.py:1-116 and dear/dopt_rsag_bo.py:100-101,148-171,317-320,401-402 —
``x`` = fusion threshold in MB, bound [1, 256], start 25; one observation per 5-iteration window
(the first window is discarded, and the first 3 timestamps of every window are dropped); the
objective is −(mean iteration time); acquisition = expected improvement with ξ = 0.1; after 10
trials jump to the best point and stop.

Differences:
  * ``bayes_opt`` is not installed: the Gaussian process (Matern 5/2, like bayes_opt) comes from
    scikit-learn and EI is maximised with random search + L-BFGS-B, both implemented here;
  * every rank measures, but only rank 0's suggestion is used and BOTH the decision and the value
    are broadcast at deterministic step counts, so ranks can never disagree about whether a
    re-bucketing happens (the reference can hang there);
  * iteration time is measured with CUDA events on the compute stream (no host synchronisation).
"""
from __future__ import annotations

import time
import warnings
from typing import Callable, List, Optional, Tuple

import numpy as np


class GaussianProcessEI:
    """1-D Bayesian optimiser: GP surrogate (Matern ν=2.5) + expected improvement."""

    def __init__(self, bound: Tuple[float, float], xi: float = 0.1, seed: int = 0):
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import Matern
        self.bound = (float(bound[0]), float(bound[1]))
        self.xi = xi
        self.rng = np.random.RandomState(seed)
        self.X: List[float] = []
        self.y: List[float] = []
        self.gp = GaussianProcessRegressor(kernel=Matern(nu=2.5), alpha=1e-6, normalize_y=True,
                                           n_restarts_optimizer=5, random_state=self.rng)

    def register(self, x: float, target: float) -> None:
        self.X.append(float(x))
        self.y.append(float(target))

    def _ei(self, xs: np.ndarray) -> np.ndarray:
        from scipy.stats import norm
        mean, std = self.gp.predict(xs.reshape(-1, 1), return_std=True)
        y_max = max(self.y)
        a = mean - y_max - self.xi
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.where(std > 0, a / std, 0.0)
        ei = a * norm.cdf(z) + std * norm.pdf(z)
        return np.where(std > 0, ei, 0.0)

    def suggest(self) -> float:
        lo, hi = self.bound
        if not self.X:
            return float(self.rng.uniform(lo, hi))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.gp.fit(np.asarray(self.X).reshape(-1, 1), np.asarray(self.y))
        from scipy.optimize import minimize
        cand = self.rng.uniform(lo, hi, size=2000)
        vals = self._ei(cand)
        best_x, best_v = float(cand[int(np.argmax(vals))]), float(np.max(vals))
        for x0 in cand[np.argsort(vals)[-5:]]:
            res = minimize(lambda x: -float(self._ei(np.asarray(x))[0]), x0=[x0], bounds=[(lo, hi)], method="L-BFGS-B")
            if res.success and -res.fun > best_v:
                best_x, best_v = float(res.x[0]), float(-res.fun)
        return float(np.clip(best_x, lo, hi))


class Tuner:
    """Tune ``x`` to minimise the iteration time (call ``step()`` once per training iteration).

    ``step()`` returns a new value of ``x`` when one should be applied, else ``None`` — on every
    rank, consistently.
    """

    def __init__(self, x: float = 25.0, bound=(1.0, 256.0), max_num_steps: int = 10, interval: int = 5,
                 rank: int = 0, broadcast: Optional[Callable] = None, clock: Optional[Callable[[], object]] = None,
                 elapsed: Optional[Callable[[object, object], float]] = None, verbose: bool = True, seed: int = 0):
        self._current_point = float(x)
        self._bound = bound
        self._max_num_steps = max_num_steps
        self._interval = interval
        self._rank = rank
        self._broadcast = broadcast or (lambda obj: obj)
        self._clock = clock or time.perf_counter
        self._elapsed = elapsed or (lambda a, b: b - a)
        self._verbose = verbose and rank == 0
        self._opt = GaussianProcessEI(bound, xi=0.1, seed=seed)
        self._opt_point: Optional[float] = None
        self._opt_iter_time: Optional[float] = None
        self._num_steps = 0
        self._marks = []
        self._warmup_record = True
        self._bo_cost: List[float] = []
        self.history: List[Tuple[float, float]] = []
        self.finished = False

    def opt_point(self):
        return self._opt_point, self._opt_iter_time

    def _record(self) -> Optional[float]:
        self._marks.append(self._clock())
        if len(self._marks) < self._interval:
            return None
        marks, self._marks = self._marks, []
        if self._warmup_record:          # the first window after a (re-)bucketing is warm-up
            self._warmup_record = False
            return None
        d = [self._elapsed(marks[i - 1], marks[i]) for i in range(3, len(marks))]
        return float(np.mean(d)) if d else None

    def step(self) -> Optional[float]:
        if self.finished:
            return None
        if self._num_steps == self._max_num_steps:
            # all ranks reach this on the same call; rank 0's optimum wins
            self.finished = True
            best = self._broadcast((self._opt_point, self._opt_iter_time))
            self._opt_point, self._opt_iter_time = best
            if self._verbose:
                print("BO Tuning optimal param: %.4f, optimal iteration time %.4f" % (best[0], best[1]))
                print("BO Tuning cost:", float(np.mean(self._bo_cost)) if self._bo_cost else 0.0)
            if self._current_point != best[0]:
                self._current_point = best[0]
                return best[0]
            return None
        iter_time = self._record()
        if iter_time is None:
            return None
        if self._verbose:
            print("BO Tuning step [%d], param: %.4f, iteration time: %.4f" % (self._num_steps, self._current_point, iter_time))
        self.history.append((self._current_point, iter_time))
        if self._opt_point is None or iter_time < self._opt_iter_time:
            self._opt_point, self._opt_iter_time = self._current_point, iter_time
        self._opt.register(self._current_point, -iter_time)
        nxt = None
        if self._rank == 0:
            t0 = time.perf_counter()
            nxt = self._opt.suggest()
            self._bo_cost.append(time.perf_counter() - t0)
        nxt = float(self._broadcast(nxt))
        self._current_point = nxt
        self._num_steps += 1
        # like the reference, the re-bucketing cost falls into the 3 dropped timestamps of the
        # next window (dear/tuner.py:66)
        return nxt


def attach_tuner(optimizer, x: float = None, bound=(1.0, 256.0), max_num_steps: int = 10, interval: int = 5,
                 verbose: bool = True) -> Tuner:
    """Enable BO tuning of the fusion threshold on a ``DistributedOptimizer`` (dopt_rsag_bo)."""
    import torch
    from .. import runtime
    eng = optimizer._dear
    if x is None:
        x = eng.threshold if eng.threshold is not None else 25.0
    cuda = eng.device.type == "cuda"
    if cuda:
        def clock():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        def elapsed(a, b):
            b.synchronize()
            return a.elapsed_time(b) / 1e3
    else:
        clock, elapsed = None, None
    tuner = Tuner(x=x, bound=bound, max_num_steps=max_num_steps, interval=interval, rank=runtime.rank(),
                  broadcast=lambda obj: runtime.broadcast_object(obj, src=0), clock=clock, elapsed=elapsed, verbose=verbose)

    def on_step():
        nxt = tuner.step()
        if nxt is not None:
            eng.threshold = nxt
            eng.request_rebucket(("threshold", float(nxt)))
    eng._step_callbacks.append(on_step)
    optimizer.tuner = tuner
    return tuner
