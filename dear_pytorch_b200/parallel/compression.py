"""Gradient compressors for the WFBP / MG-WFBP baselines.

Same registry and call shapes as the reference (``*/compression.py:258-267``):
``compress(tensor, name=None, sigma_scale=.., ratio=..) -> (tensor, indexes, values)`` for the
sparsifiers, ``(packed, None, None)`` for the sign compressors; ``decompress``; ``add_residuals``.
The DeAR path itself never compresses (the reference only passes ``--compressor none`` through,
dear/imagenet_benchmark.py:18,52,114).

Unlike the reference, the sign compressors are functional: the missing external ``bit2byte``
extension (dear/compression.py:110,137) is replaced by a bit-packing written with torch integer ops.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch


class NoneCompressor:
    name = "none"

    def compress(self, tensor, name=None, **_):
        return tensor, tensor.dtype

    def decompress(self, tensor, ctc=None):
        return tensor

    def clear(self):
        pass


class TopKCompressor:
    """Magnitude top-k sparsification; the unsent part is kept as a residual (Aji & Heafield 2017)."""
    name = "topk"
    error_feedback = False

    def __init__(self):
        self.residuals: Dict[object, torch.Tensor] = {}
        self.values: Dict[object, torch.Tensor] = {}
        self.indexes: Dict[object, torch.Tensor] = {}
        self.zero_conditions: Dict[object, torch.Tensor] = {}
        self.zc: Optional[torch.Tensor] = None
        self.current_ratio = 1.0

    def clear(self):
        self.residuals.clear()
        self.values.clear()
        self.indexes.clear()
        self.zero_conditions.clear()

    def get_residuals(self, name, like_tensor):
        if name not in self.residuals:
            self.residuals[name] = torch.zeros_like(like_tensor)
        return self.residuals[name]

    def _select(self, flat: torch.Tensor, k: int, sigma_scale: float) -> torch.Tensor:
        return torch.topk(flat.abs(), k=k, sorted=False)[1]

    @torch.no_grad()
    def compress(self, tensor, name=None, sigma_scale=2.5, ratio=0.05):
        flat = tensor.view(-1)
        res = self.get_residuals(name, flat)
        k = max(int(flat.numel() * ratio), 1)
        self.current_ratio = ratio
        if self.error_feedback:
            flat.add_(res)
        idx = self._select(flat, k, sigma_scale)
        vals = flat[idx]
        res.copy_(flat)
        res[idx] = 0.0
        zc = self.zero_conditions.get(name)
        if zc is None or zc.numel() != flat.numel():
            zc = self.zero_conditions[name] = torch.ones_like(flat, dtype=torch.float32)
        zc.fill_(1.0)
        zc[idx] = 0.0
        self.zc = zc
        if idx.numel() < k:
            # data-dependent selectors (gaussian) may return fewer than k entries: every rank must contribute the
            # SAME number to the all-gather, so pad with (index 0, value 0), which adds nothing when scattered
            pad = k - idx.numel()
            idx = torch.cat([idx, idx.new_zeros(pad)])
            vals = torch.cat([vals, vals.new_zeros(pad)])
        self.values[name], self.indexes[name] = vals, idx
        return tensor, idx, vals

    @torch.no_grad()
    def add_residuals(self, included_indexes, name):
        """Put back into the residual the selected values that did NOT make the global cut."""
        vals = self.values[name].clone()
        if not torch.is_tensor(included_indexes):
            included_indexes = torch.as_tensor(included_indexes, device=vals.device)
        vals[included_indexes.long()] = 0.0
        self.residuals[name][self.indexes[name]] += vals

    def decompress(self, tensor, original_tensor_size=None):
        return tensor


class EFTopKCompressor(TopKCompressor):
    name = "eftopk"
    error_feedback = True


class GTopKCompressor(TopKCompressor):
    """Local top-k selection for the gTop-k sparse all-reduce (Shi et al., ICDCS 2019): the optimizer exchanges the
    selections pairwise in log2(P) rounds (``baselines.gtopk``) and calls ``add_residuals`` so that locally selected
    values that did not survive the global cut go back into the residual.  The reference keys this path off the
    compressor's name (``'gtopk' in name``, wfbp/dopt.py:725-726) but never registers such a compressor."""
    name = "gtopk"
    error_feedback = False


class EFGTopKCompressor(GTopKCompressor):
    name = "gtopkef"
    error_feedback = True


class GaussianCompressor(TopKCompressor):
    """Threshold from a normal fit of the gradient (Shi et al. 2019), refined in <= 3 rounds."""
    name = "gaussian"
    error_feedback = True

    def _select(self, flat, k, sigma_scale):
        mean, std = float(flat.mean()), float(flat.std())
        ratio = k / flat.numel()
        # two-sided tail of N(mean, std) holding a `ratio` fraction of the mass
        z = math.sqrt(2.0) * _erfinv(1.0 - ratio)
        thres = abs(mean) + z * std
        a = flat.abs()
        idx = (a > thres).nonzero().view(-1)
        for _ in range(3):
            if idx.numel() < 2 * k / 3:
                thres *= 0.5
            elif idx.numel() > 4 * k / 3:
                thres *= 1.5
            else:
                break
            idx = (a > thres).nonzero().view(-1)
        if idx.numel() == 0:
            idx = torch.topk(a, k=1)[1]
        return idx[:k]


def _erfinv(x: float) -> float:
    return float(torch.erfinv(torch.tensor(x, dtype=torch.float64)))


class SignCompressor:
    """signSGD with majority vote: 1 bit per element, 32 elements per int32 word."""
    name = "signum"

    def __init__(self):
        self.zc = None
        self.residuals: Dict[object, torch.Tensor] = {}

    def clear(self):
        self.residuals.clear()

    @staticmethod
    def packing(src: torch.Tensor):
        sign = torch.sign(src)
        bits = (sign.view(-1) >= 0).to(torch.int64)          # 1 = non-negative
        pad = (-bits.numel()) % 32
        if pad:
            bits = torch.cat([bits, bits.new_ones(pad)])
        w = (bits.view(-1, 32) << torch.arange(32, device=bits.device)).sum(1)
        w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
        return w, sign

    @staticmethod
    def unpacking(words: torch.Tensor, size) -> torch.Tensor:
        n = 1
        for s in size:
            n *= int(s)
        w = words.to(torch.int64) & 0xFFFFFFFF
        bits = (w.view(-1, 1) >> torch.arange(32, device=words.device)) & 1
        return (bits.view(-1)[:n].float() * 2.0 - 1.0).view(*size)

    @classmethod
    def majority_vote(cls, word_list):
        n = word_list[0].numel() * 32
        votes = sum(cls.unpacking(w, (n,)) for w in word_list)
        return cls.packing(votes)[0]

    def _before(self, name, tensor):
        pass

    def _after(self, name, sign, tensor):
        pass

    @torch.no_grad()
    def compress(self, tensor, name=None, sigma_scale=3, ratio=0.05):
        self._before(name, tensor)
        words, sign = self.packing(tensor)
        self._after(name, sign, tensor)
        return words, None, None

    def decompress(self, tensor, original_tensor_size):
        return self.unpacking(tensor, original_tensor_size)


class EFSignCompressor(SignCompressor):
    name = "efsignum"

    def _before(self, name, tensor):
        if name not in self.residuals:
            self.residuals[name] = torch.zeros_like(tensor)
        tensor.add_(self.residuals[name])

    def _after(self, name, sign, tensor):
        self.residuals[name] = tensor - sign.view_as(tensor)


compressors = {
    "none": NoneCompressor,
    None: NoneCompressor,
    "topk": TopKCompressor,
    "eftopk": EFTopKCompressor,
    "gaussian": GaussianCompressor,
    "gtopk": GTopKCompressor,
    "gtopkef": EFGTopKCompressor,
    "signum": SignCompressor,
    "efsignum": EFSignCompressor,
}
