"""shared helpers of the test-suite (no product code, no oracle code)"""
import numpy as np


def _mix(z):
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def row_hashes(num, jtag, jimage):
    """order-independent 64-bit hash of every neighbor row: sum_j mix(tag_j*32 + image_j)"""
    h = _mix(jtag.astype(np.uint64) * np.uint64(32) + jimage.astype(np.uint64))
    out = np.zeros(len(num), np.uint64)
    rows = np.repeat(np.arange(len(num)), num)
    with np.errstate(over="ignore"):
        np.add.at(out, rows, h)
    return out


def relerr(a, b):
    """max |a-b| / max |b|  (the tolerance definition used by every parity test)"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return np.inf
    if a.size == 0:
        return 0.0
    s = np.abs(b).max()
    d = np.abs(a - b).max()
    return d / s if s > 0 else d


def relerr_elem(a, b, floor=1e-4):
    """element-wise figure: max_i |a_i - b_i| / max(|b_i|, floor * max|b|).  Components that cancel to (almost) nothing -- the
    force on a particle at rest inside the bulk -- are measured against the floor; everything above it against its own size."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return np.inf
    if a.size == 0:
        return 0.0
    s = np.abs(b).max()
    if not s > 0:
        return float(np.abs(a - b).max())
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor * s)).max())
