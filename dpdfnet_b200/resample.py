"""Batched streaming sample-rate conversion on the device.

The reference converts every chunk on the host with a stateless ``librosa.resample(..., res_type="soxr_hq")`` call
(``audio.py:20-27``, used by ``stream.py:112,163-164`` and ``api.py:86,110``).  librosa / soxr are not available
offline, so *parity with that filter is unpinned*; what is pinned here is the published polyphase algorithm of
``scipy.signal.resample_poly`` (Kaiser-windowed sinc, ``half_len = 10 * max(up, down)``), which the host helper
``audio.ensure_sample_rate`` already uses.  ``BatchResampler`` computes exactly that sum for B streams at once and
keeps per-stream history, so chunked streaming output equals the one-shot result (no chunk-boundary artefacts).
"""
from __future__ import annotations

import ctypes
from fractions import Fraction
from typing import Optional

import numpy as np

from .engine import _ptr, _raise, load_library


def design_taps(up: int, down: int) -> np.ndarray:
    """The zero-phase FIR of ``scipy.signal.resample_poly(x, up, down)`` (default Kaiser window, beta 5), float32."""
    from scipy.signal import firwin
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    return np.ascontiguousarray(h, dtype=np.float32)


class BatchResampler:
    """``sr_in -> sr_out`` for up to ``max_streams`` streams advancing in lock step (device tensors in and out)."""

    def __init__(self, sr_in: int, sr_out: int, max_streams: int, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("BatchResampler needs a CUDA device: there is no CPU fallback (use audio.ensure_sample_rate)")
        frac = Fraction(int(sr_out), int(sr_in))
        self.up, self.down = frac.numerator, frac.denominator
        self.max_streams, self.device = int(max_streams), int(device)
        self._lib = load_library()
        self._taps = design_taps(self.up, self.down)
        self._h = ctypes.c_void_p()
        rc = self._lib.dpdf_resampler_create(self.up, self.down, self._taps.ctypes.data, self._taps.size, self.max_streams,
                                             self.device, ctypes.byref(self._h))
        if rc:
            _raise(self._lib, rc)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.dpdf_resampler_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def reset(self):
        import torch
        rc = self._lib.dpdf_resampler_reset(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            _raise(self._lib, rc)

    def process(self, chunk, flush: bool = False):
        """chunk: float32 CUDA tensor [B, n] (n may be 0 with ``flush``) -> [B, n_out] of the samples that are now final."""
        import torch
        B, n = chunk.shape
        assert chunk.is_cuda and chunk.dtype == torch.float32 and (n == 0 or chunk.stride(1) == 1)
        n_out = int(self._lib.dpdf_resampler_pending(self._h, n, int(flush)))
        if n_out < 0:
            raise ValueError("bad chunk size")
        out = torch.empty((B, n_out), device=chunk.device, dtype=torch.float32)
        got = ctypes.c_int64()
        rc = self._lib.dpdf_resampler_process(self._h, _ptr(chunk) if n else None, chunk.stride(0) if n else 0, n,
                                              _ptr(out) if n_out else None, out.stride(0) if n_out else 0, B, int(flush),
                                              ctypes.byref(got), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            _raise(self._lib, rc)
        assert got.value == n_out
        return out

    def resample(self, x):
        """One-shot: [B, n] -> [B, ceil(n * up / down)], equal to ``scipy.signal.resample_poly`` row by row."""
        self.reset()
        return self.process(x, flush=True)
