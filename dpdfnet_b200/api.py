"""Whole-clip enhancement (mirror of ``package/src/dpdfnet/api.py``).

``enhance`` keeps the reference signature and the reference's alignment convention (pad one window,
centred STFT, per-frame runtime calls, attenuation limit against the 4-frame-delayed noisy spectrum,
centred iSTFT, drop ``2*win``; ``api.py:51-113``).  File I/O, downloads and the CLI are out of scope.
``enhance_batch`` is the batched, device-resident variant the engine exists for.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import numpy as np

from .models import DEFAULT_MODEL, available_model_entries, resolve_model


def available_models() -> List[Dict[str, Any]]:
    return available_model_entries()


def _enhance_with_runtime(audio: np.ndarray, sample_rate: int, *, runtime, model_sample_rate: int,
                          attn_limit_db: Optional[float] = None,
                          progress_callback: Optional[Callable[[int, int], None]] = None) -> np.ndarray:
    from .audio import (apply_attn_limit, ensure_sample_rate, fit_length, make_stft_config, postprocess_spec,
                        preprocess_waveform, to_mono)
    from .onnx_backend import infer_win_len

    wave = to_mono(np.asarray(audio, dtype=np.float32))
    sr_in = int(sample_rate)
    x = ensure_sample_rate(wave, sr_in, model_sample_rate)
    cfg = make_stft_config(infer_win_len(runtime.session, model_sample_rate))
    spec = preprocess_waveform(np.pad(x, (0, cfg.win_len)), cfg)          # alignment padding, api.py:88
    T = int(spec.shape[1])
    state = runtime.init_state.copy()
    outs: List[np.ndarray] = []
    if progress_callback is not None:
        progress_callback(0, T)
    for t in range(T):
        frame = np.ascontiguousarray(spec[:, t:t + 1], dtype=np.float32)
        y, state = runtime.session.run([runtime.out_spec_name, runtime.out_state_name],
                                       {runtime.in_spec_name: frame, runtime.in_state_name: state})
        outs.append(np.ascontiguousarray(y, dtype=np.float32))
        if progress_callback is not None:
            progress_callback(t + 1, T)
    if not outs:
        return wave.copy()
    spec_e = apply_attn_limit(spec, np.concatenate(outs, axis=1), attn_limit_db)
    y = ensure_sample_rate(postprocess_spec(spec_e, cfg), model_sample_rate, sr_in)
    return fit_length(y, wave.shape[0]).astype(np.float32, copy=False)


def enhance(audio: np.ndarray, sample_rate: int, *, model: str = DEFAULT_MODEL,
            onnx_path: Optional[Union[str, Path]] = None, attn_limit_db: Optional[float] = None,
            verbose: bool = False, progress_callback: Optional[Callable[[int, int], None]] = None) -> np.ndarray:
    from .onnx_backend import build_runtime_model
    resolved = resolve_model(model=model, onnx_path=onnx_path, auto_download=True, verbose=verbose)
    runtime = build_runtime_model(resolved.onnx_path)
    return _enhance_with_runtime(audio, sample_rate, runtime=runtime, model_sample_rate=resolved.info.sample_rate,
                                 attn_limit_db=attn_limit_db, progress_callback=progress_callback)


def _batch_resample(waves: List[np.ndarray], sr_in: int, sr_out: int, device: int) -> List[np.ndarray]:
    """All clips through the device resampler in one call (zero-extended to the longest: a polyphase FIR sees zeros
    past the end of a clip either way, so every clip gets exactly its own ``resample_poly`` result)."""
    if sr_in == sr_out or not waves:
        return waves
    import torch
    from .resample import BatchResampler
    n = max(w.size for w in waves)
    if n == 0:
        return waves
    x = np.zeros((len(waves), n), np.float32)
    for i, w in enumerate(waves):
        x[i, :w.size] = w
    rs = BatchResampler(sr_in, sr_out, max_streams=len(waves), device=device)
    y = rs.resample(torch.from_numpy(x).to(f"cuda:{device}")).cpu().numpy()
    rs.close()
    return [y[i, :-(-w.size * rs.up // rs.down)].copy() for i, w in enumerate(waves)]


def fit_to(y: np.ndarray, n: int) -> np.ndarray:
    return y[:n] if y.size >= n else np.pad(y, (0, n - y.size))


def enhance_batch(clips: Sequence[np.ndarray], sample_rate: int, *, model: str = DEFAULT_MODEL,
                  onnx_path: Optional[Union[str, Path]] = None, engine=None, exact_offline: bool = True,
                  attn_limit_db: Optional[float] = None) -> List[np.ndarray]:
    """Enhance many mono clips at once on one engine (streams = clips, ragged lengths allowed).

    With ``exact_offline`` the result equals the reference's offline PyTorch model (``model/dpdfnet.py``) on every
    clip: each stream follows its own reflect padding and flush schedule (``offline.py:enhance_offline_exact_ragged``),
    so a short clip batched with long ones gets the same samples it gets alone.  ``attn_limit_db`` limits the
    attenuation like the reference (``audio.py:50-76``: ``alpha * noisy + (1 - alpha) * enhanced``, alpha =
    10^(-dB/20)); the offline-exact output is sample-aligned with its input and the STFT pair reconstructs exactly, so
    the blend is done on the waveform.

    Clips of ``hop`` samples or fewer cannot be reflect-padded by the offline model (``torch.stft(center=True)`` raises
    on them, model/modules.py:360-369); they come back as silence of the input length, with a ``UserWarning``.
    """
    from .audio import _validate_attn_limit_db, ensure_sample_rate, to_mono
    from .offline import enhance_offline_exact_ragged
    from .onnx_backend import create_session
    resolved = resolve_model(model=model, onnx_path=onnx_path)
    if not len(clips):
        return []
    if engine is None:
        engine = create_session(resolved.onnx_path, max_streams=len(clips)).engine
    model_sr = resolved.info.sample_rate
    waves = _batch_resample([to_mono(np.asarray(c, dtype=np.float32)) for c in clips], int(sample_rate), model_sr, engine.device)
    if not exact_offline:
        raise NotImplementedError("only the offline-exact schedule is implemented for batches")
    attn_limit_db = _validate_attn_limit_db(attn_limit_db)                    # NaN / negative: ValueError (audio.py:41-47)
    hop = engine.spec.hop
    short = [i for i, w in enumerate(waves) if w.size <= hop]
    if short:
        import warnings
        warnings.warn(f"{len(short)} clip(s) of <= {hop} samples are too short for the offline model and are returned as silence")
    outs = enhance_offline_exact_ragged(engine, [w for w in waves if w.size > hop])
    it = iter(outs)
    res = []
    for i, w in enumerate(waves):
        y = np.zeros(w.size, np.float32)
        if i not in short:
            e = next(it)
            y[:e.size] = e
            if attn_limit_db is not None:
                alpha = np.float32(10.0 ** (-float(attn_limit_db) / 20.0))
                y[:e.size] = alpha * w[:e.size] + (np.float32(1.0) - alpha) * y[:e.size]
        res.append(y)
    res = _batch_resample(res, model_sr, int(sample_rate), engine.device)
    return [np.ascontiguousarray(fit_to(y, np.asarray(clips[i]).shape[0]), dtype=np.float32) for i, y in enumerate(res)]
