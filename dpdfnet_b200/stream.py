"""Chunked real-time enhancement (mirror of ``package/src/dpdfnet/stream.py``).

Same public behaviour as the reference ``StreamEnhancer``: arbitrary chunk sizes, no output before
one full window has arrived, then ``hop`` samples per completed frame, ``flush()`` zero-pads the
remainder, ``reset()`` starts a new stream, a sample-rate change raises ``ValueError``.

Two execution paths share that behaviour:

* engine path - when the runtime's session is an :class:`~dpdfnet_b200.onnx_backend.EngineSession`
  the windowing, real DFT, network, inverse DFT and overlap-add of every hop run fused on the GPU
  (``dpdf_prime_pcm`` / ``dpdf_run_pcm_host``); only PCM crosses the bus and the recurrent state
  never leaves the device.
* session path - any other object with the ONNX-Runtime call shape (the reference's seam,
  ``stream.py:129-135``) is driven frame by frame with host DSP, exactly like the reference.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Union

import numpy as np

from .audio import ensure_sample_rate, make_stft_config, to_mono
from .models import DEFAULT_MODEL, resolve_model
from .onnx_backend import EngineSession, RuntimeModel, build_runtime_model, infer_win_len

_EMPTY = np.zeros(0, dtype=np.float32)


class StreamEnhancer:
    def __init__(self, model: str = DEFAULT_MODEL, onnx_path: Optional[Union[str, Path]] = None,
                 verbose: bool = False) -> None:
        resolved = resolve_model(model=model, onnx_path=onnx_path, auto_download=True, verbose=verbose)
        self._runtime: RuntimeModel = build_runtime_model(resolved.onnx_path)
        self._model_sr: int = resolved.info.sample_rate
        self._win_len: int = infer_win_len(self._runtime.session, self._model_sr)
        cfg = make_stft_config(self._win_len)
        self._hop_size: int = cfg.hop_size
        self._window: np.ndarray = cfg.window
        self._freq_bins: int = self._win_len // 2 + 1
        sess = self._runtime.session
        self._fused = isinstance(sess, EngineSession) and sess.engine.spec.win == self._win_len
        self._input_sr: Optional[int] = None
        self.reset()

    # ------------------------------------------------------------------
    def reset(self) -> None:
        """Forget the stream: recurrent state, analysis history and overlap-add tail."""
        self._state: np.ndarray = self._runtime.init_state.copy()
        self._in_buf: np.ndarray = _EMPTY
        self._out_buf: np.ndarray = np.zeros(self._win_len, dtype=np.float32)
        self._input_sr = None
        self._primed = False
        if self._fused:
            s = self._runtime.session
            s.engine.reset([s.slot])

    def process(self, chunk: np.ndarray, sample_rate: Optional[int] = None) -> np.ndarray:
        x = to_mono(np.asarray(chunk, dtype=np.float32))
        if x.size == 0:
            return _EMPTY.copy()
        sr = self._model_sr if sample_rate is None else sample_rate
        if self._input_sr is None:
            self._input_sr = sr
        elif sr != self._input_sr:
            raise ValueError(f"Sample rate changed from {self._input_sr} to {sr} between process() calls.  "
                             "Call reset() before processing a new stream.")
        self._in_buf = np.concatenate([self._in_buf, ensure_sample_rate(x, sr, self._model_sr)])
        out = self._run_fused() if self._fused else self._run_session()
        if out.size and sr != self._model_sr:
            return ensure_sample_rate(out, self._model_sr, sr)
        return out

    def flush(self) -> np.ndarray:
        """Zero-pad what is buffered to one more window and return at most one hop of audio."""
        pending = self._in_buf.size + (self._hop_size if self._primed else 0)
        if pending == 0:
            return _EMPTY.copy()
        sr = self._input_sr or self._model_sr
        out = self.process(np.zeros(self._win_len - pending, dtype=np.float32), sample_rate=self._model_sr)
        out = out[:self._hop_size]
        if sr != self._model_sr:
            out = ensure_sample_rate(out, self._model_sr, sr)
        return out.astype(np.float32)

    # ----- engine path ----------------------------------------------------
    def _run_fused(self) -> np.ndarray:
        sess = self._runtime.session
        hop = self._hop_size
        if not self._primed:
            if self._in_buf.size < self._win_len:       # the reference emits nothing before one full window
                return _EMPTY.copy()
            sess.engine.prime_pcm_host(self._in_buf[None, :hop], slot_ids=[sess.slot])
            self._in_buf = self._in_buf[hop:]
            self._primed = True
        T = self._in_buf.size // hop
        if T == 0:
            return _EMPTY.copy()
        out = sess.engine.run_pcm_host(self._in_buf[None, :T * hop], slot_ids=[sess.slot])[0]
        self._in_buf = self._in_buf[T * hop:]
        return out

    # ----- generic session path (reference seam) ----------------------------
    def _run_session(self) -> np.ndarray:
        rt, win, hop = self._runtime, self._win_len, self._hop_size
        done: List[np.ndarray] = []
        while self._in_buf.size >= win:
            X = np.fft.rfft(self._in_buf[:win] * self._window, n=win)
            spec = np.stack([X.real, X.imag], axis=-1).astype(np.float32)[None, None]
            spec_e, self._state = rt.session.run([rt.out_spec_name, rt.out_state_name],
                                                 {rt.in_spec_name: spec, rt.in_state_name: self._state})
            y = np.asarray(spec_e)[0, 0]
            frame = (np.fft.irfft(y[:, 0] + 1j * y[:, 1], n=win) * self._window).astype(np.float32)
            acc = self._out_buf + frame
            done.append(acc[:hop].copy())                 # Vorbis COLA: first hop is final after this frame
            self._out_buf = np.concatenate([acc[hop:], np.zeros(hop, dtype=np.float32)])
            self._in_buf = self._in_buf[hop:]
        # keep the reference's accounting: the retained hop of context lives in _in_buf here
        return np.concatenate(done) if done else _EMPTY.copy()
