"""Small helpers of the reference's ``utils.py`` (dear/utils.py:8-59) that its optimizers and plotting scripts import:
run ids, output directories, timer dictionaries, a numpy top-k, the Gaussian-k threshold scale and two matplotlib
conveniences (duck-typed on the axes / bar objects, matplotlib itself is not needed to import this module)."""
from __future__ import annotations

import hashlib
import os
import time
from typing import Dict, List

import numpy as np


def gen_random_id() -> str:
    """Hex digest unique to this call (the reference hashes ``str(time.time())`` without encoding it: py2-only)."""
    return hashlib.sha256(("%r-%d-%d" % (time.time(), os.getpid(), time.perf_counter_ns())).encode()).hexdigest()


def create_path(relative_path: str, base: str = None) -> str:
    """``mkdir -p`` relative to ``base`` (default: the current directory; the reference resolves against its source
    directory, which is read-only for an installed package).  Returns the absolute path."""
    path = os.path.join(base or os.getcwd(), relative_path)
    os.makedirs(path, exist_ok=True)
    return os.path.abspath(path)


def force_insert_item(d: Dict, key, val) -> None:
    """Append ``val`` to the list under ``key`` (the timers of the WFBP optimizers: name -> [seconds, ...])."""
    d.setdefault(key, []).append(val)


def topk(tensor: np.ndarray, k: int):
    """Indices and values of the ``k`` largest-magnitude entries of a 1-D numpy array (unordered)."""
    k = min(int(k), tensor.size)
    idx = np.argpartition(np.abs(tensor), tensor.size - k)[tensor.size - k:] if k > 0 else np.empty(0, dtype=np.int64)
    return idx, tensor[idx]


def get_approximate_sigma_scale(density: float) -> float:
    """How many standard deviations the Gaussian-k threshold starts from for a target density
    (> 0.7: 0.5;  0.05-0.7: 1.5;  0.01-0.05: 2.0;  below: 3.0)."""
    for bound, scale in ((0.7, 0.5), (0.05, 1.5), (0.01, 2.0)):
        if density > bound:
            return scale
    return 3.0


def update_fontsize(ax, fontsize: float = 12.0) -> None:
    """Set one font size on the title, axis labels and tick labels of a matplotlib axes."""
    for item in [ax.title, ax.xaxis.label, ax.yaxis.label, *ax.get_xticklabels(), *ax.get_yticklabels()]:
        item.set_fontsize(fontsize)


def autolabel(rects: List, ax, label: str, rotation: float = 90) -> None:
    """Write ``label`` just above every bar of a bar plot."""
    for r in rects:
        top = r.get_y() + r.get_height()
        ax.text(r.get_x() + r.get_width() / 2.0, 1.03 * top, label, ha="center", va="bottom", rotation=rotation)
