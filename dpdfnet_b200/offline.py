"""Whole-clip enhancement on the streaming engine with the *offline-exact* schedule.

The shipped ``enhance()`` path (``api.py:88-111``) is the naive stream and matches the offline
PyTorch model (``model/dpdfnet.py:540-570``) only after a warm-up transient.  The schedule below
reproduces the offline model over the whole clip:

1. frame like ``torch.stft(center=True, pad_mode='reflect')`` (``model/modules.py:357-370``);
2. for the first ``conv_lookahead`` = 2 frames advance only the normalisers and the spectral delay
   rings (``pad_feat`` crops those feature frames, ``model/dpdfnet.py:463,545-546``) -> WARMUP flag;
3. after the last real frame run 4 flush hops with a zero spectrum, the first 2 of them with zero
   features (``model/dpdfnet.py:463``, ``model/multiframe.py:74``) -> ZERO_SPEC / ZERO_FEAT flags;
4. output frame of hop tau is offline frame tau-4; overlap-add hops 1..T-1 are the waveform.
"""
from __future__ import annotations

import numpy as np

from .engine import FLAG_WARMUP, FLAG_ZERO_FEAT, FLAG_ZERO_SPEC, Engine


def enhance_offline_exact(engine: Engine, wave: np.ndarray) -> np.ndarray:
    """wave [B, n] float32 (host) -> enhanced [B, hop*(T-1)], T = 1 + n // hop."""
    import torch
    sp = engine.spec
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    if wave.ndim != 2:
        raise ValueError(f"expected [B, n] audio, got {wave.shape}")
    B, n = wave.shape
    hop = sp.hop
    if n <= hop:
        raise ValueError("clip shorter than one hop + 1 sample cannot be reflect-padded")
    if B > engine.max_streams:
        raise ValueError(f"B={B} exceeds max_streams={engine.max_streams}")
    T = 1 + n // hop
    dev = f"cuda:{engine.device}"
    x = torch.from_numpy(wave).to(dev)
    padded = torch.cat([x[:, 1:hop + 1].flip(1), x, x[:, n - hop - 1:n - 1].flip(1),
                        torch.zeros(B, 4 * hop, device=dev)], 1).contiguous()
    engine.reset(list(range(B)))
    engine.prime_pcm(padded[:, :hop])
    out = torch.zeros(B, (T + 4) * hop, device=dev)
    warm = torch.full((B,), FLAG_WARMUP, dtype=torch.int32, device=dev)
    engine.run_pcm(padded[:, hop:3 * hop], flags=warm, out=out[:, :2 * hop])
    if T > 2:
        engine.run_pcm(padded[:, 3 * hop:(T + 1) * hop], out=out[:, 2 * hop:T * hop])
    fl = torch.full((B,), FLAG_ZERO_SPEC | FLAG_ZERO_FEAT, dtype=torch.int32, device=dev)
    engine.run_pcm(padded[:, (T + 1) * hop:(T + 3) * hop], flags=fl, out=out[:, T * hop:(T + 2) * hop])
    fl2 = torch.full((B,), FLAG_ZERO_SPEC, dtype=torch.int32, device=dev)
    engine.run_pcm(padded[:, (T + 3) * hop:(T + 5) * hop], flags=fl2, out=out[:, (T + 2) * hop:(T + 4) * hop])
    return out[:, 5 * hop:(T + 4) * hop].cpu().numpy()
