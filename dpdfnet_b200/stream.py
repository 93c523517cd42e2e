"""Chunked real-time enhancement (mirror of ``package/src/dpdfnet/stream.py``).

Same public behaviour as the reference ``StreamEnhancer``: arbitrary chunk sizes, no output before
one full window has arrived, then ``hop`` samples per completed frame, ``flush()`` zero-pads the
remainder, ``reset()`` starts a new stream, a sample-rate change raises ``ValueError``.

Two execution paths share that behaviour:

* engine path - when the runtime's session is an :class:`~dpdfnet_b200.onnx_backend.EngineSession`
  the windowing, real DFT, network, inverse DFT and overlap-add of every hop run fused on the GPU
  (``dpdf_prime_pcm`` / ``dpdf_run_pcm_host``); only PCM crosses the bus and the recurrent state
  never leaves the device.  The session is a *slot of the model's shared engine*
  (:class:`~dpdfnet_b200.onnx_backend.EnginePool`), so any number of enhancers cost one engine and
  :func:`process_many` sends the ready hops of many enhancers through ONE batched ``dpdf_step_pcm`` call.
  When the caller's sample rate differs from the model's, the conversion is a *stateful* polyphase
  resampler on the device (``resample.py:BatchResampler``, one in each direction, per stream): chunked
  output equals one-shot output, where the reference's stateless per-chunk ``librosa.resample``
  (``stream.py:112,163-164``) restarts its filter at every chunk edge.
* session path - any other object with the ONNX-Runtime call shape (the reference's seam,
  ``stream.py:129-135``) is driven frame by frame with host DSP, exactly like the reference.

:class:`StreamGroup` is the engine-native form of the same API for B streams that advance in lock step
(fixed-size packets from B callers): one object, ``[B, n]`` arrays in and out, everything between the two
PCIe copies on the device.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .audio import ensure_sample_rate, make_stft_config, to_mono
from .models import DEFAULT_MODEL, resolve_model
from .onnx_backend import EnginePool, EngineSession, RuntimeModel, build_runtime_model, infer_win_len

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
        self._rs_in = self._rs_out = None       # device resamplers of the engine path (created on first use)
        self.reset()

    # ------------------------------------------------------------------
    def reset(self) -> None:
        """Forget the stream: recurrent state, analysis history, overlap-add tail and resampler history."""
        self._state: np.ndarray = self._runtime.init_state.copy()
        self._in_buf: np.ndarray = _EMPTY
        self._out_buf: np.ndarray = np.zeros(self._win_len, dtype=np.float32)
        self._input_sr = None
        self._primed = False
        if self._fused:
            s = self._runtime.session
            with s.lock:
                s.engine.reset([s.slot])
        for rs in (self._rs_in, self._rs_out):
            if rs is not None:
                rs.close()
        self._rs_in = self._rs_out = None

    def close(self) -> None:
        """Return the engine slot to the shared pool (the reference has no counterpart: its session dies with the object)."""
        for rs in (self._rs_in, self._rs_out):
            if rs is not None:
                rs.close()
        self._rs_in = self._rs_out = None
        sess = getattr(self._runtime, "session", None)
        if isinstance(sess, EngineSession):
            sess.close()
        self._fused = False

    # ----- sample-rate conversion around the path ---------------------------------------------------
    def _resample(self, which: str, x: np.ndarray, flush: bool = False) -> np.ndarray:
        """Engine path: stateful device resampler (`which` = "in": caller rate -> model rate, "out": back)."""
        import torch
        from .resample import BatchResampler
        sess = self._runtime.session
        rs = self._rs_in if which == "in" else self._rs_out
        if rs is None:
            a, b = (self._input_sr, self._model_sr) if which == "in" else (self._model_sr, self._input_sr)
            rs = BatchResampler(a, b, max_streams=1, device=sess.engine.device)
            if which == "in":
                self._rs_in = rs
            else:
                self._rs_out = rs
        dev = f"cuda:{sess.engine.device}"
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)[None]).to(dev) if x.size else torch.empty((1, 0), device=dev)
        return rs.process(t, flush=flush)[0].cpu().numpy()

    def _ingest(self, chunk: np.ndarray, sample_rate: Optional[int]) -> Tuple[bool, int]:
        """Validate and buffer a chunk at the model rate.  Returns (anything buffered, caller's rate)."""
        x = to_mono(np.asarray(chunk, dtype=np.float32))
        if x.size == 0:
            return False, self._input_sr or self._model_sr
        sr = self._model_sr if sample_rate is None else sample_rate
        if self._input_sr is None:
            self._input_sr = sr
        elif sr != self._input_sr:
            raise ValueError(f"Sample rate changed from {self._input_sr} to {sr} between process() calls.  "
                             "Call reset() before processing a new stream.")
        if sr == self._model_sr:
            y = x
        elif self._fused:
            y = self._resample("in", x)
        else:
            y = ensure_sample_rate(x, sr, self._model_sr)
        self._in_buf = np.concatenate([self._in_buf, y])
        return True, sr

    def _emit(self, out: np.ndarray, sr: int) -> np.ndarray:
        if sr == self._model_sr:
            return out
        if self._fused:
            return self._resample("out", out)
        return ensure_sample_rate(out, self._model_sr, sr) if out.size else out

    def process(self, chunk: np.ndarray, sample_rate: Optional[int] = None) -> np.ndarray:
        got, sr = self._ingest(chunk, sample_rate)
        if not got:
            return _EMPTY.copy()
        out = self._run_fused() if self._fused else self._run_session()
        return self._emit(out, sr)

    def flush(self) -> np.ndarray:
        """Zero-pad what is buffered to one more window and return the audio it completes (at most one hop at the
        model rate; with the engine path's stateful resamplers also the few samples their filters still held)."""
        sr = self._input_sr or self._model_sr
        stateful = self._fused and sr != self._model_sr
        head = _EMPTY
        if stateful and self._rs_in is not None:
            self._in_buf = np.concatenate([self._in_buf, self._resample("in", _EMPTY, flush=True)])
            if self._in_buf.size >= self._win_len or (self._primed and self._in_buf.size >= self._hop_size):
                head = self._run_fused()                     # the filter tail completed whole frames
        pending = self._in_buf.size + (self._hop_size if self._primed else 0)
        if pending == 0 and not (stateful and self._rs_out is not None):
            return _EMPTY.copy()
        out = _EMPTY
        if pending:
            self._in_buf = np.concatenate([self._in_buf, np.zeros(self._win_len - pending, dtype=np.float32)])
            out = (self._run_fused() if self._fused else self._run_session())[:self._hop_size]
        if stateful:
            out = np.concatenate([head, out])
            return np.concatenate([self._resample("out", out), self._resample("out", _EMPTY, flush=True)]).astype(np.float32)
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
            with sess.lock:
                sess.engine.prime_pcm_host(self._in_buf[None, :hop], slot_ids=[sess.slot])
            self._in_buf = self._in_buf[hop:]
            self._primed = True
        T = self._in_buf.size // hop
        if T == 0:
            return _EMPTY.copy()
        with sess.lock:
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


# =====================================================================================================
# Many enhancers, one batched call
# =====================================================================================================
def process_many(enhancers: Sequence[StreamEnhancer], chunks: Sequence[np.ndarray],
                 sample_rate: Optional[int] = None) -> List[np.ndarray]:
    """``[e.process(c, sample_rate) for e, c in zip(enhancers, chunks)]`` with the engine work batched.

    Every enhancer buffers its chunk exactly as ``process`` does; then, per shared engine, the enhancers that have the
    same number T of complete hops pending go through ONE ``dpdf_run_pcm_host`` call (``[n, T*hop]`` rows, their slot
    ids), instead of one B=1 call each.  With equal-sized packets - the real-time server case - that is a single
    call for all of them.  Results are identical to the one-by-one calls (streams never interact; the engine's
    batched step is row-independent).  Enhancers on the generic session path are processed one by one.
    """
    if len(enhancers) != len(chunks):
        raise ValueError("enhancers and chunks must have the same length")
    if len({id(e) for e in enhancers}) != len(enhancers):
        raise ValueError("an enhancer may appear only once per call")
    rates: List[int] = []
    ready: List[bool] = []
    for e, c in zip(enhancers, chunks):
        got, sr = e._ingest(c, sample_rate)
        ready.append(got)
        rates.append(sr)
    outs: List[np.ndarray] = [_EMPTY.copy() for _ in enhancers]
    groups: Dict[int, List[int]] = {}
    for i, e in enumerate(enhancers):
        if not ready[i]:
            continue
        if not e._fused:
            outs[i] = e._emit(e._run_session(), rates[i])
        else:
            groups.setdefault(id(e._runtime.session.engine), []).append(i)
    for idx in groups.values():
        _run_fused_many([enhancers[i] for i in idx], idx, outs)
        for i in idx:
            outs[i] = enhancers[i]._emit(outs[i], rates[i])
    return outs


def _run_fused_many(es: Sequence[StreamEnhancer], idx: Sequence[int], outs: List[np.ndarray]) -> None:
    """The engine part of ``process_many`` for enhancers that share one engine."""
    sess0 = es[0]._runtime.session
    eng, hop, win = sess0.engine, es[0]._hop_size, es[0]._win_len
    with sess0.lock:
        first = [e for e in es if not e._primed and e._in_buf.size >= win]
        if first:                      # the reference emits nothing before one full window (stream.py:116)
            eng.prime_pcm_host(np.stack([e._in_buf[:hop] for e in first]), slot_ids=[e._runtime.session.slot for e in first])
            for e in first:
                e._in_buf = e._in_buf[hop:]
                e._primed = True
        by_T: Dict[int, List[int]] = {}
        for k, e in enumerate(es):
            T = e._in_buf.size // hop if e._primed else 0
            if T:
                by_T.setdefault(T, []).append(k)
        for T, ks in by_T.items():
            pcm = np.stack([es[k]._in_buf[:T * hop] for k in ks])
            y = eng.run_pcm_host(pcm, slot_ids=[es[k]._runtime.session.slot for k in ks])
            for r, k in enumerate(ks):
                es[k]._in_buf = es[k]._in_buf[T * hop:]
                outs[idx[k]] = y[r]


def flush_many(enhancers: Sequence[StreamEnhancer]) -> List[np.ndarray]:
    """``[e.flush() for e in enhancers]``; the final frames of enhancers on the model rate are batched."""
    plain = [e for e in enhancers if e._fused and (e._input_sr or e._model_sr) == e._model_sr]
    res: Dict[int, np.ndarray] = {}
    pads, who = [], []
    for e in plain:
        pending = e._in_buf.size + (e._hop_size if e._primed else 0)
        if pending == 0:
            res[id(e)] = _EMPTY.copy()
        else:
            pads.append(np.zeros(e._win_len - pending, dtype=np.float32))
            who.append(e)
    if who:
        for e, y in zip(who, process_many(who, pads, sample_rate=None)):
            res[id(e)] = y[:e._hop_size].astype(np.float32)
    return [res[id(e)] if id(e) in res else e.flush() for e in enhancers]


class StreamGroup:
    """B streams of one model advancing in lock step: the ``StreamEnhancer`` contract for ``[B, n]`` arrays.

    Row b of every call behaves like its own ``StreamEnhancer`` (same buffering, same first-window latency, same
    ``flush`` accounting - ``tests/test_gpu_stream_pool.py`` checks it row by row), but the group holds B slots of the
    shared engine and the whole path between the host->device copy of the packet and the device->host copy of the
    result runs on the device: stateful batched resampling when ``sample_rate`` differs from the model's
    (``resample.py``), a device FIFO for samples that do not fill a hop yet, ``dpdf_prime_pcm`` for the first window,
    one ``dpdf_run_pcm`` for all complete hops.  ``process`` accepts and returns numpy arrays (pinned staging inside)
    or CUDA tensors (no PCIe copies at all).
    """

    def __init__(self, model: str = DEFAULT_MODEL, streams: int = 1, onnx_path: Optional[Union[str, Path]] = None,
                 device: int = 0, verbose: bool = False) -> None:
        import torch
        if streams <= 0:
            raise ValueError("streams must be positive")
        resolved = resolve_model(model=model, onnx_path=onnx_path, auto_download=True, verbose=verbose)
        self._pool = EnginePool.get(resolved.onnx_path, device, first=streams)
        self.engine, self.slots = self._pool.acquire(streams)
        self._model_sr = resolved.info.sample_rate
        self._hop, self._win = self.engine.spec.hop, self.engine.spec.win
        self.streams = int(streams)
        self._dev = torch.device(f"cuda:{self.engine.device}")
        contiguous = self.slots == list(range(self.slots[0], self.slots[0] + streams)) and self.slots[0] == 0
        self._slots_dev = None if contiguous else torch.tensor(self.slots, dtype=torch.int32, device=self._dev)
        self._fifo = torch.zeros((streams, 4 * self._win), device=self._dev)
        self._pin_in = self._pin_out = None
        self._rs_in = self._rs_out = None
        self._input_sr: Optional[int] = None
        self._fill = 0
        self._primed = False

    # ------------------------------------------------------------------
    def close(self) -> None:
        for rs in (self._rs_in, self._rs_out):
            if rs is not None:
                rs.close()
        self._rs_in = self._rs_out = None
        if self.engine is not None:
            for s in self.slots:
                self._pool.release(self.engine, s)
            self.engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        with self._pool.lock:
            self.engine.reset(self.slots)
        for rs in (self._rs_in, self._rs_out):
            if rs is not None:
                rs.reset()
        self._input_sr = None
        self._fill = 0
        self._primed = False

    # ------------------------------------------------------------------
    def _to_device(self, x):
        import torch
        if isinstance(x, torch.Tensor):
            return x.to(self._dev, torch.float32), True
        a = np.asarray(x, dtype=np.float32)
        if a.ndim != 2 or a.shape[0] != self.streams:
            raise ValueError(f"chunks must be [{self.streams}, n], got {a.shape}")
        if self._pin_in is None or self._pin_in.shape[1] < a.shape[1]:
            self._pin_in = torch.empty((self.streams, max(a.shape[1], self._hop)), dtype=torch.float32).pin_memory()
        view = self._pin_in[:, :a.shape[1]]
        view.numpy()[...] = a
        return view.to(self._dev, non_blocking=True), False

    def _to_caller(self, y, as_tensor: bool):
        import torch
        if as_tensor:
            return y
        n = y.shape[1]
        if n == 0:
            torch.cuda.current_stream(self._dev).synchronize()      # the pinned input block is reused by the next call
            return np.zeros((self.streams, 0), np.float32)
        if self._pin_out is None or self._pin_out.shape[1] < n:
            self._pin_out = torch.empty((self.streams, n), dtype=torch.float32).pin_memory()
        view = self._pin_out[:, :n]
        view.copy_(y, non_blocking=True)
        torch.cuda.current_stream(self._dev).synchronize()
        return view.numpy().copy()

    def _append(self, y) -> None:
        import torch
        n = y.shape[1]
        if self._fill + n > self._fifo.shape[1]:
            grown = torch.zeros((self.streams, 2 * (self._fill + n)), device=self._dev)
            grown[:, :self._fill] = self._fifo[:, :self._fill]
            self._fifo = grown
        self._fifo[:, self._fill:self._fill + n] = y
        self._fill += n

    def _consume(self, n: int) -> None:
        rest = self._fill - n
        if rest:
            self._fifo[:, :rest] = self._fifo[:, n:self._fill].clone()
        self._fill = rest

    def _run(self):
        """Prime on the first full window, then every complete hop in one engine call -> [B, T*hop] device tensor."""
        import torch
        hop = self._hop
        with self._pool.lock:
            if not self._primed:
                if self._fill < self._win:
                    return torch.empty((self.streams, 0), device=self._dev)
                self.engine.prime_pcm(self._fifo[:, :hop], slot_ids=self._slots_dev)
                self._consume(hop)
                self._primed = True
            T = self._fill // hop
            if T == 0:
                return torch.empty((self.streams, 0), device=self._dev)
            out = self.engine.run_pcm(self._fifo[:, :T * hop], slot_ids=self._slots_dev)
        self._consume(T * hop)
        return out

    def process(self, chunks, sample_rate: Optional[int] = None):
        """chunks ``[B, n]`` (every stream gets n new samples) -> ``[B, m]`` enhanced samples at ``sample_rate``."""
        import torch
        with torch.cuda.device(self._dev):
            x, as_tensor = self._to_device(chunks)
            if x.shape[1] == 0:
                return self._to_caller(torch.empty((self.streams, 0), device=self._dev), as_tensor)
            sr = self._model_sr if sample_rate is None else int(sample_rate)
            if self._input_sr is None:
                self._input_sr = sr
            elif sr != self._input_sr:
                raise ValueError(f"Sample rate changed from {self._input_sr} to {sr} between process() calls.  "
                                 "Call reset() before processing a new stream.")
            if sr != self._model_sr:
                from .resample import BatchResampler
                if self._rs_in is None:
                    self._rs_in = BatchResampler(sr, self._model_sr, self.streams, device=self.engine.device)
                    self._rs_out = BatchResampler(self._model_sr, sr, self.streams, device=self.engine.device)
                x = self._rs_in.process(x.contiguous())
            self._append(x)
            out = self._run()
            if sr != self._model_sr and out.shape[1]:
                out = self._rs_out.process(out.contiguous())
            return self._to_caller(out, as_tensor)

    def flush(self, as_tensor: bool = False):
        """Zero-pad every stream to one more window; returns what that completes (``StreamEnhancer.flush`` per row)."""
        import torch
        with torch.cuda.device(self._dev):
            sr = self._input_sr or self._model_sr
            head = torch.empty((self.streams, 0), device=self._dev)
            if sr != self._model_sr and self._rs_in is not None:
                self._append(self._rs_in.process(torch.empty((self.streams, 0), device=self._dev), flush=True))
                head = self._run()
            pending = self._fill + (self._hop if self._primed else 0)
            out = torch.empty((self.streams, 0), device=self._dev)
            if pending:
                self._append(torch.zeros((self.streams, self._win - pending), device=self._dev))
                out = self._run()[:, :self._hop]
            if sr != self._model_sr and self._rs_out is not None:
                out = torch.cat([head, out], 1)
                out = torch.cat([self._rs_out.process(out.contiguous()) if out.shape[1] else out,
                                 self._rs_out.process(torch.empty((self.streams, 0), device=self._dev), flush=True)], 1)
            return self._to_caller(out, as_tensor)
