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


def enhance_offline_exact_ragged(engine: Engine, waves) -> list:
    """Clips of different lengths in one batched run, each with exactly the result ``enhance_offline_exact`` gives it
    alone (the bulk ``enhance-dir`` use case, ``cli.py:220-326``).

    Every stream follows its own schedule -- its own reflect padding, its own flush hops -- through per-hop, per-stream
    flags; hops on which all flag rows agree are submitted as one multi-hop run.  Returns a list of 1-D arrays,
    ``hop * (T_i - 1)`` samples each."""
    import torch
    sp = engine.spec
    hop = sp.hop
    waves = [np.ascontiguousarray(w, dtype=np.float32).reshape(-1) for w in waves]
    B = len(waves)
    if B == 0:
        return []
    if B > engine.max_streams:
        raise ValueError(f"B={B} exceeds max_streams={engine.max_streams}")
    if min(w.size for w in waves) <= hop:
        raise ValueError("clip shorter than one hop + 1 sample cannot be reflect-padded")
    T = [1 + w.size // hop for w in waves]
    Tm = max(T)
    pcm = np.zeros((B, (Tm + 5) * hop), np.float32)
    flags = np.full((Tm + 4, B), FLAG_ZERO_SPEC, np.int32)           # after a stream's last flush hop: output ignored
    for i, w in enumerate(waves):
        n = w.size
        pcm[i, :hop] = w[1:hop + 1][::-1]
        pcm[i, hop:hop + n] = w
        pcm[i, hop + n:2 * hop + n] = w[n - hop - 1:n - 1][::-1]
        flags[:2, i] = FLAG_WARMUP
        flags[2:T[i], i] = 0
        flags[T[i]:T[i] + 2, i] = FLAG_ZERO_SPEC | FLAG_ZERO_FEAT
        flags[T[i] + 2:T[i] + 4, i] = FLAG_ZERO_SPEC
    dev = f"cuda:{engine.device}"
    x = torch.from_numpy(pcm).to(dev)
    fl = torch.from_numpy(flags).to(dev)
    out = torch.zeros(B, (Tm + 4) * hop, device=dev)
    engine.reset(list(range(B)))
    engine.prime_pcm(x[:, :hop])
    t0 = 0
    while t0 < Tm + 4:                                               # maximal runs of hops with identical flag rows
        t1 = t0 + 1
        while t1 < Tm + 4 and np.array_equal(flags[t1], flags[t0]):
            t1 += 1
        engine.run_pcm(x[:, (t0 + 1) * hop:(t1 + 1) * hop], flags=fl[t0], out=out[:, t0 * hop:t1 * hop])
        t0 = t1
    res = out.cpu().numpy()
    return [res[i, 5 * hop:(T[i] + 4) * hop].copy() for i in range(B)]
