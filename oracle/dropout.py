"""Restatement of the library's counter-based dropout masks (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference trains with torch's dropout (DROPOUT_P, models/transformer.py:38-44, positional_encoding.py:35); its
masks come from torch's RNG stream and cannot be matched by another implementation, so libsdvg draws its own masks
from a hash of (seed, step, site, element) - csrc/common.cuh: fmix32 / drop_hash / drop_site_key - and this module
restates that hash with torch integer ops so that the oracle can run the reference's arithmetic under exactly the
same masks (tests/test_gpu_train.py)."""
import numpy as np
import torch

M32 = 0xFFFFFFFF
SA_P, SA_OUT, CA_P, CA_OUT, FF_H, FF_OUT = range(6)


def fmix32(h):
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & M32
    return h ^ (h >> 16)


def site_key(seed, step, site):
    inner = fmix32((site * 0x632BE5AB + step) & M32)
    return fmix32((seed & M32) ^ inner) ^ ((seed >> 32) & M32)


def enc_site(layer, which):
    return 1000 + layer * 10 + which


def dec_site(layer, which):
    return 2000 + layer * 10 + which


class Dropper:
    """drop(site, x2d): x2d (rows, cols) -> x2d * mask / (1 - p) with the library's mask for that site and step."""

    def __init__(self, p, seed, step):
        self.p32 = np.float32(p)
        self.seed, self.step = int(seed), int(step)
        t = float(self.p32) * 4294967296.0
        self.thr = 4294967295 if t >= 4294967295.0 else max(int(t), 1)
        self.scale = float(np.float32(1.0) / (np.float32(1.0) - self.p32))

    def mask(self, site, rows, cols):
        idx = (torch.arange(rows, dtype=torch.int64)[:, None] * cols + torch.arange(cols, dtype=torch.int64)[None, :]) & M32
        key = site_key(self.seed, self.step, site)
        h = fmix32((idx * 0x9E3779B1 + key) & M32)
        return h >= self.thr

    def __call__(self, site, x2d):
        if float(self.p32) <= 0.0:
            return x2d
        keep = self.mask(site, x2d.shape[0], x2d.shape[1])
        return x2d * (keep.to(x2d.dtype) * self.scale)
