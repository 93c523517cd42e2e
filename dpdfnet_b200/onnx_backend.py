"""The runtime seam, B200 edition (mirror of ``package/src/dpdfnet/onnx_backend.py``).

``build_runtime_model`` returns the same ``RuntimeModel`` shape the reference's callers consume
(``api.py:98-101``, ``stream.py:129-135``), but ``session`` is an :class:`EngineSession` that runs the
frame on the GPU engine through the C ABI instead of ONNX Runtime's CPU provider.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import weights as _weights
from .engine import Engine
from .models import MODEL_REGISTRY, RANDOM_WEIGHTS
from .spec import NB_DF, get_spec


class _IO:
    def __init__(self, name: str, shape):
        self.name, self.shape, self.type = name, list(shape), "tensor(float)"


class EngineSession:
    """``onnxruntime.InferenceSession``-shaped view of one engine slot.

    ``run([spec_e, state_out], {spec: f32[1,1,F,2], state_in: f32[S]})`` keeps the reference call shape
    (the caller owns and round-trips the flat state), so it works for B=1 drop-in use.  High-throughput
    callers should talk to ``self.engine`` directly (batched, state resident on the device).
    """

    def __init__(self, engine: Engine, slot: int = 0, pool: Optional["EnginePool"] = None):
        self.engine, self.slot, self._pool = engine, int(slot), pool
        F, S = engine.spec.freq_bins, engine.spec.state_size
        self._inputs = [_IO("spec", (1, 1, F, 2)), _IO("state_in", (S,))]
        self._outputs = [_IO("spec_e", (1, 1, F, 2)), _IO("state_out", (S,))]
        self._resident = None          # state array object this slot currently mirrors

    def close(self) -> None:
        """Give the slot back to the shared engine (pooled sessions only); the session is unusable afterwards."""
        pool, self._pool = self._pool, None
        if pool is not None:
            pool.release(self.engine, self.slot)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def lock(self):
        """Engine handles are single-writer (include/dpdfnet_b200.h): sessions sharing an engine serialise on this."""
        return self._pool.lock if self._pool is not None else _NO_LOCK

    def get_inputs(self) -> List[_IO]:
        return self._inputs

    def get_outputs(self) -> List[_IO]:
        return self._outputs

    def get_providers(self) -> List[str]:
        return ["DPDFNetB200ExecutionProvider"]

    def run(self, output_names: Sequence[str], feed: Dict[str, np.ndarray]):
        spec = np.ascontiguousarray(feed["spec"], dtype=np.float32)
        state = np.asarray(feed["state_in"], dtype=np.float32)
        F = self.engine.spec.freq_bins
        if spec.shape != (1, 1, F, 2):
            raise ValueError(f"spec must have shape (1, 1, {F}, 2), got {spec.shape}")
        if state.ndim != 1 or state.size != self.engine.spec.state_size:
            raise ValueError(f"state size mismatch: expected {self.engine.spec.state_size}, got {state.shape}")
        with self.lock:
            if state is not self._resident:      # caller handed back the array we produced: already on device
                self.engine.state_import(self.slot, state)
            out = self.engine.step_spec_host(spec.reshape(1, F, 2), slot_ids=[self.slot])
            new_state = self.engine.state_export(self.slot)
        self._resident = new_state
        res = {"spec_e": out.reshape(1, 1, F, 2), "state_out": new_state}
        return [res[n] for n in (output_names or ["spec_e", "state_out"])]


@dataclass(frozen=True)
class RuntimeModel:
    session: Any
    init_state: np.ndarray
    in_spec_name: str
    in_state_name: str
    out_spec_name: str
    out_state_name: str


def _model_name_from_path(path: Path) -> str:
    stem = path.name if path.parent == RANDOM_WEIGHTS else path.stem
    if stem not in MODEL_REGISTRY:
        raise ValueError(f"Cannot infer the model name from '{path.name}'; expected one of {', '.join(MODEL_REGISTRY)}")
    return stem


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_LOCK = _NoLock()


def _load_weights(path: Path):
    """(spec, checkpoint) for a weights path: reference ``.pth`` state_dict, packed ``.dpdfw`` blob, a reference
    ``.onnx`` export (initialisers + metadata, onnx_ingest.py) or the seeded random stand-in."""
    if path.suffix == ".onnx" and path.is_file():
        from .onnx_ingest import load_onnx_checkpoint
        return load_onnx_checkpoint(path)
    spec = get_spec(_model_name_from_path(path))
    if path.parent == RANDOM_WEIGHTS:
        return spec, None
    if not path.is_file():
        raise FileNotFoundError(f"Model weights file not found: {path}")
    if path.suffix == ".dpdfw":
        return spec, path.read_bytes()
    return spec, _weights.load_checkpoint_file(path)


def _make_engine(spec, ckpt, max_streams: int, device: int) -> Engine:
    try:
        return Engine(spec, ckpt, max_streams=max_streams, device=device)
    except (ValueError, FileNotFoundError):
        raise
    except Exception as exc:  # noqa: BLE001
        raise RuntimeError("Failed to initialise the DPDFNet B200 engine (no CPU fallback exists).") from exc


class EnginePool:
    """Engines shared by every streaming session of one (weights, device) pair.

    The reference runs one ORT session per ``StreamEnhancer`` (stream.py:41-47, README.md:153-154); here every
    enhancer of a model takes a *slot* of a shared batched engine, so the ready hops of many enhancers can go through
    one ``dpdf_step_pcm`` call (``stream.process_many`` / ``StreamGroup``).  The pool starts with one engine of
    ``$DPDFNET_B200_POOL_STREAMS`` slots (default 128, or what ``reserve`` asked for) and adds engines of twice the
    previous size when it runs out."""

    _pools: Dict[Tuple[str, int], "EnginePool"] = {}
    _registry_lock = threading.Lock()

    def __init__(self, path: Path, device: int, first: int):
        self.path, self.device = path, int(device)
        self.spec, self._ckpt = _load_weights(path)
        if not isinstance(self._ckpt, (bytes, bytearray)) :
            ck = self._ckpt if self._ckpt is not None else _weights.random_checkpoint(self.spec, 0)
            sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in ck.items()}
            self._ckpt = _weights.pack_checkpoint(self.spec, sd)          # pack once, reuse for every engine
        self.lock = threading.RLock()
        self.engines: List[Engine] = []
        self._free: List[List[int]] = []
        self._next = max(1, int(first))

    @classmethod
    def get(cls, path: Union[str, Path], device: int = 0, first: Optional[int] = None) -> "EnginePool":
        key = (str(path), int(device))
        with cls._registry_lock:
            pool = cls._pools.get(key)
            if pool is None:
                n = first if first is not None else int(os.environ.get("DPDFNET_B200_POOL_STREAMS", "128"))
                pool = cls._pools[key] = EnginePool(Path(path), device, n)
            return pool

    @classmethod
    def shutdown(cls) -> None:
        """Destroy every pooled engine (tests; a server calls it at exit)."""
        with cls._registry_lock:
            for pool in cls._pools.values():
                for e in pool.engines:
                    e.close()
            cls._pools.clear()

    @property
    def capacity(self) -> int:
        return sum(e.max_streams for e in self.engines)

    @property
    def in_use(self) -> int:
        return self.capacity - sum(len(f) for f in self._free)

    def _grow(self, at_least: int) -> None:
        n = max(self._next, at_least)
        self.engines.append(_make_engine(self.spec, self._ckpt, n, self.device))
        self._free.append(list(range(n - 1, -1, -1)))
        self._next = 2 * n

    def acquire(self, n: int = 1) -> Tuple[Engine, List[int]]:
        """`n` slots of ONE engine (so they can be stepped together), freshly reset."""
        with self.lock:
            idx = next((i for i, f in enumerate(self._free) if len(f) >= n), None)
            if idx is None:
                self._grow(n)
                idx = len(self.engines) - 1
            slots = [self._free[idx].pop() for _ in range(n)]
            self.engines[idx].reset(slots)
            return self.engines[idx], slots

    def release(self, engine: Engine, slot: int) -> None:
        with self.lock:
            for e, f in zip(self.engines, self._free):
                if e is engine and slot not in f:
                    f.append(int(slot))


def reserve(model: str, streams: int, onnx_path: Optional[Union[str, Path]] = None, device: int = 0) -> EnginePool:
    """Size the shared engine of `model` for `streams` concurrent enhancers before creating them."""
    from .models import resolve_model
    pool = EnginePool.get(resolve_model(model=model, onnx_path=onnx_path).onnx_path, device, first=streams)
    with pool.lock:
        free = max((len(f) for f in pool._free), default=0)
        if free < streams:
            pool._grow(streams)
    return pool


def create_session(weights_path: Union[str, Path], max_streams: Optional[int] = None, device: int = 0) -> EngineSession:
    """Counterpart of ``create_cpu_session`` (onnx_backend.py:21-49).  Without ``max_streams`` the session is one slot
    of the shared engine of that model (:class:`EnginePool`); with it, a private engine of that many slots."""
    path = Path(weights_path)
    if max_streams is None:
        if path.suffix != ".onnx":
            _model_name_from_path(path)                               # ValueError for unknown names before any GPU work
        if path.parent != RANDOM_WEIGHTS and not path.is_file():
            raise FileNotFoundError(f"Model weights file not found: {path}")
        pool = EnginePool.get(path, device)
        engine, slots = pool.acquire(1)
        return EngineSession(engine, slots[0], pool)
    spec, ckpt = _load_weights(path)
    return EngineSession(_make_engine(spec, ckpt, max_streams, device), 0)


def initial_state(spec) -> np.ndarray:
    """[mu0, s0, zeros...] - what ``load_initial_state_from_metadata`` rebuilds (onnx_backend.py:52-78)."""
    mu0, s0 = _weights.norm_init(spec)
    st = np.zeros(spec.state_size, dtype=np.float32)
    st[:spec.fe_feat] = mu0
    st[spec.fe_feat:spec.fe_feat + NB_DF] = s0
    return st


def build_runtime_model(onnx_path: Union[str, Path]) -> RuntimeModel:
    """Mirror of onnx_backend.py:81-99.  The session is a slot of the model's shared engine, so any number of
    ``StreamEnhancer`` / ``enhance`` callers cost one engine, not one each."""
    session = create_session(onnx_path, device=int(os.environ.get("DPDFNET_B200_DEVICE", "0")))
    ins, outs = session.get_inputs(), session.get_outputs()
    if Path(onnx_path).suffix == ".onnx":          # the shipped artefact: state init from its metadata, like the reference (:52-78)
        from .onnx_ingest import initial_state_from_metadata, read_onnx
        init = initial_state_from_metadata(read_onnx(onnx_path))
        if init.size != session.engine.spec.state_size:
            raise ValueError(f"ONNX metadata state_size {init.size} does not match the engine's {session.engine.spec.state_size}")
    else:
        init = initial_state(session.engine.spec)
    return RuntimeModel(session=session, init_state=init, in_spec_name=ins[0].name,
                        in_state_name=ins[1].name, out_spec_name=outs[0].name, out_state_name=outs[1].name)


def infer_win_len(session: Any, default_sr: int) -> int:
    """(F - 1) * 2 from the first input's shape, else 20 ms (onnx_backend.py:102-107)."""
    shape = session.get_inputs()[0].shape
    bins = shape[-2] if len(shape) >= 2 else None
    if isinstance(bins, int) and bins > 1:
        return (bins - 1) * 2
    return int(round(default_sr * 0.02))
