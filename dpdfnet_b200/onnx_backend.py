"""The runtime seam, B200 edition (mirror of ``package/src/dpdfnet/onnx_backend.py``).

``build_runtime_model`` returns the same ``RuntimeModel`` shape the reference's callers consume
(``api.py:98-101``, ``stream.py:129-135``), but ``session`` is an :class:`EngineSession` that runs the
frame on the GPU engine through the C ABI instead of ONNX Runtime's CPU provider.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Sequence, Union

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

    def __init__(self, engine: Engine, slot: int = 0):
        self.engine, self.slot = engine, int(slot)
        F, S = engine.spec.freq_bins, engine.spec.state_size
        self._inputs = [_IO("spec", (1, 1, F, 2)), _IO("state_in", (S,))]
        self._outputs = [_IO("spec_e", (1, 1, F, 2)), _IO("state_out", (S,))]
        self._resident = None          # state array object this slot currently mirrors

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
        if state is not self._resident:          # caller handed back the array we produced: already on device
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


def create_session(weights_path: Union[str, Path], max_streams: int = 1, device: int = 0) -> EngineSession:
    """Counterpart of ``create_cpu_session`` (onnx_backend.py:21-49): builds the GPU engine."""
    path = Path(weights_path)
    name = _model_name_from_path(path)
    spec = get_spec(name)
    if path.parent == RANDOM_WEIGHTS:
        ckpt = None
    elif not path.is_file():
        raise FileNotFoundError(f"Model weights file not found: {path}")
    elif path.suffix == ".dpdfw":
        ckpt = path.read_bytes()
    else:
        ckpt = _weights.load_checkpoint_file(path)
    try:
        engine = Engine(spec, ckpt, max_streams=max_streams, device=device)
    except (ValueError, FileNotFoundError):
        raise
    except Exception as exc:  # noqa: BLE001
        raise RuntimeError("Failed to initialise the DPDFNet B200 engine (no CPU fallback exists).") from exc
    return EngineSession(engine, 0)


def initial_state(spec) -> np.ndarray:
    """[mu0, s0, zeros...] - what ``load_initial_state_from_metadata`` rebuilds (onnx_backend.py:52-78)."""
    mu0, s0 = _weights.norm_init(spec)
    st = np.zeros(spec.state_size, dtype=np.float32)
    st[:spec.fe_feat] = mu0
    st[spec.fe_feat:spec.fe_feat + NB_DF] = s0
    return st


def build_runtime_model(onnx_path: Union[str, Path]) -> RuntimeModel:
    n = int(os.environ.get("DPDFNET_B200_MAX_STREAMS", "1"))
    session = create_session(onnx_path, max_streams=n, device=int(os.environ.get("DPDFNET_B200_DEVICE", "0")))
    ins, outs = session.get_inputs(), session.get_outputs()
    return RuntimeModel(session=session, init_state=initial_state(session.engine.spec), in_spec_name=ins[0].name,
                        in_state_name=ins[1].name, out_spec_name=outs[0].name, out_state_name=outs[1].name)


def infer_win_len(session: Any, default_sr: int) -> int:
    """(F - 1) * 2 from the first input's shape, else 20 ms (onnx_backend.py:102-107)."""
    shape = session.get_inputs()[0].shape
    bins = shape[-2] if len(shape) >= 2 else None
    if isinstance(bins, int) and bins > 1:
        return (bins - 1) * 2
    return int(round(default_sr * 0.02))
