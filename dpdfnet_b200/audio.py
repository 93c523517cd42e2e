"""Host-side DSP helpers with the reference's names and semantics (``package/src/dpdfnet/audio.py``).

Only numpy (+ scipy for resampling) is needed.  librosa is not a dependency: the centred STFT /
iSTFT pair below restates ``librosa.stft(center=True, pad_mode='reflect')`` and
``librosa.istft(center=True)`` as used at ``audio.py:104-136``.  Resampling is a polyphase filter
(``scipy.signal.resample_poly``) instead of librosa's soxr_hq - parity for that step is unpinned
(SURVEY.md section 8f, rank 1).
"""
from __future__ import annotations

from dataclasses import dataclass
from fractions import Fraction
from typing import Optional

import numpy as np

ATTN_LIMIT_NOISY_FRAME_OFFSET = 4        # mask look-ahead 2 + deep-filter look-ahead 2 (audio.py:8)


def to_mono(audio: np.ndarray) -> np.ndarray:
    """1-D audio is returned as float32; [n, channels] is averaged over channels (audio.py:11-17)."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 2:
        return a.mean(axis=1, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError(f"Expected mono/stereo audio, got shape {a.shape}")
    return a


def ensure_sample_rate(audio: np.ndarray, sample_rate: int, target_sample_rate: int) -> np.ndarray:
    a = np.asarray(audio, dtype=np.float32)
    if int(sample_rate) == int(target_sample_rate) or a.size == 0:
        return a
    from scipy.signal import resample_poly
    frac = Fraction(int(target_sample_rate), int(sample_rate))
    return resample_poly(a.astype(np.float64), frac.numerator, frac.denominator).astype(np.float32)


def fit_length(audio: np.ndarray, target_len: int) -> np.ndarray:
    a = np.asarray(audio, dtype=np.float32).reshape(-1)
    n = a.shape[0]
    if n >= target_len:
        return a[:target_len]
    return np.concatenate([a, np.zeros(target_len - n, dtype=np.float32)])


def _validate_attn_limit_db(attn_limit_db: Optional[float]) -> Optional[float]:
    if attn_limit_db is None:
        return None
    v = float(attn_limit_db)
    if v != v or v < 0.0:
        raise ValueError("attn_limit_db must be non-negative, infinity, or None.")
    return v


def apply_attn_limit(spec_noisy: np.ndarray, spec_enh: np.ndarray, attn_limit_db: Optional[float]) -> np.ndarray:
    """``alpha * noisy[t-4] + (1 - alpha) * enhanced[t]`` with ``alpha = 10^(-dB/20)`` (audio.py:50-76)."""
    limit = _validate_attn_limit_db(attn_limit_db)
    enh = np.asarray(spec_enh, dtype=np.float32)
    if limit is None:
        return enh
    noisy = np.asarray(spec_noisy, dtype=np.float32)
    if noisy.shape != enh.shape:
        raise ValueError(f"spec_noisy and spec_enh must have matching shapes, got {noisy.shape} and {enh.shape}.")
    d = ATTN_LIMIT_NOISY_FRAME_OFFSET
    delayed = np.zeros_like(noisy)
    if noisy.shape[1] > d:
        delayed[:, d:] = noisy[:, :noisy.shape[1] - d]
    alpha = float(10.0 ** (-limit / 20.0))
    return np.ascontiguousarray(alpha * delayed + (1.0 - alpha) * enh, dtype=np.float32)


def pcm16_safe(audio: np.ndarray) -> np.ndarray:
    return (np.clip(np.asarray(audio, dtype=np.float32), -1.0, 1.0) * 32767.0).astype(np.int16)


def vorbis_window(window_len: int) -> np.ndarray:
    n = np.arange(window_len, dtype=np.float64) + 0.5
    return np.sin(0.5 * np.pi * np.sin(np.pi * n / window_len) ** 2).astype(np.float32)


@dataclass(frozen=True)
class StftConfig:
    win_len: int
    hop_size: int
    window: np.ndarray


def make_stft_config(win_len: int) -> StftConfig:
    return StftConfig(win_len=int(win_len), hop_size=int(win_len) // 2, window=vorbis_window(int(win_len)))


def preprocess_waveform(waveform: np.ndarray, cfg: StftConfig) -> np.ndarray:
    """Centred, reflect-padded STFT -> float32 [1, T, F, 2], T = 1 + len // hop (audio.py:104-117)."""
    x = np.asarray(waveform, dtype=np.float32).reshape(-1)
    half = cfg.win_len // 2
    if x.size <= half:
        raise ValueError("waveform too short for reflect padding")
    padded = np.pad(x, (half, half), mode="reflect")
    T = 1 + x.size // cfg.hop_size
    idx = np.arange(T)[:, None] * cfg.hop_size + np.arange(cfg.win_len)[None, :]
    spec = np.fft.rfft(padded[idx] * cfg.window[None, :], axis=1).astype(np.complex64)
    return np.stack([spec.real, spec.imag], axis=-1).astype(np.float32)[None]


def postprocess_spec(spec_e: np.ndarray, cfg: StftConfig) -> np.ndarray:
    """Centred iSTFT, then drop the first 2*win samples and zero-pad the tail (audio.py:120-136)."""
    s = np.asarray(spec_e[0], dtype=np.float32)
    frames = np.fft.irfft(s[..., 0] + 1j * s[..., 1], n=cfg.win_len, axis=1) * cfg.window[None, :]
    T = frames.shape[0]
    total = cfg.win_len + cfg.hop_size * (T - 1)
    y = np.zeros(total, dtype=np.float64)
    env = np.zeros(total, dtype=np.float64)
    w2 = cfg.window.astype(np.float64) ** 2
    for t in range(T):
        a = t * cfg.hop_size
        y[a:a + cfg.win_len] += frames[t]
        env[a:a + cfg.win_len] += w2
    ok = env > np.finfo(np.float32).tiny
    y[ok] /= env[ok]
    half = cfg.win_len // 2
    wave = y[half:total - half].astype(np.float32)
    drop = cfg.win_len * 2
    out = np.zeros_like(wave)
    keep = max(wave.size - drop, 0)
    out[:keep] = wave[drop:drop + keep]
    return out
