"""CPU restatement of the polyphase sum the device resampler evaluates (TEST INFRASTRUCTURE ONLY: imported by tests/).

    y[m] = sum_j x[j] * h[half + m*down - j*up],   m = 0 .. ceil(n*up/down) - 1

with the zero-phase FIR ``h`` of ``dpdfnet_b200.resample.design_taps`` -- the arithmetic of
``scipy.signal.resample_poly`` / ``upfirdn`` (scipy/signal/_signaltools.py: resample_poly; the reference itself calls
librosa.resample(res_type="soxr_hq"), audio.py:20-27, which is absent offline: parity with *that* filter is unpinned,
this restatement is pinned against scipy's published algorithm in tests/test_oracle_resample.py).
"""
from __future__ import annotations

import numpy as np


def resample_direct(x: np.ndarray, up: int, down: int, taps: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    h = np.asarray(taps, dtype=np.float64)
    half = h.size // 2
    n = x.size
    n_out = -(-n * up // down)
    y = np.zeros(n_out)
    for m in range(n_out):
        c = m * down
        j_lo = max(0, -((half - c) // up) if c - half < 0 else (c - half + up - 1) // up)
        j_hi = min(n - 1, (c + half) // up)
        if j_hi >= j_lo:
            j = np.arange(j_lo, j_hi + 1)
            y[m] = np.dot(x[j], h[half + c - j * up])
    return y


def streamed_count(n_total: int, up: int, down: int, half: int, flush: bool) -> int:
    """Outputs that are final after n_total input samples (dpdf_resampler_pending)."""
    cap = -(-n_total * up // down)
    if flush:
        return cap
    lim = (n_total - 1) * up - half
    return 0 if lim < 0 else min(cap, lim // down + 1)
