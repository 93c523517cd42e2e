"""Model registry and local weight resolution (mirror of ``package/src/dpdfnet/models.py``).

Only what the hot path needs is kept: the six model names with their sample rates
(``models.py:26-69``) and a resolver that finds a *local* checkpoint.  Hugging Face download,
retries, file locks and cache directories are out of scope.

Weights are looked up, in order, in the explicit ``onnx_path`` argument (kept under its reference
name; a reference ``.pth`` state_dict or a packed ``.dpdfw`` blob is expected), then in
``$DPDFNET_MODEL_DIR`` as ``<name>.pth`` / ``checkpoints/<name>.pth`` / ``<name>.dpdfw``.  With
``DPDFNET_B200_RANDOM_WEIGHTS=1`` a seeded random checkpoint is used when nothing is found (there
are no shipped weights offline); otherwise ``FileNotFoundError`` is raised like the reference does
for a missing ONNX file (``onnx_backend.py:23-24``).
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Union

from .spec import MODEL_SPECS


@dataclass(frozen=True)
class ModelInfo:
    name: str
    sample_rate: int
    frame_ms: float
    description: str
    onnx_filename: str


_DESCRIPTIONS = {
    "baseline": "Fastest and lowest-compute baseline model.",
    "dpdfnet2": "Balanced quality/speed DPDFNet-2 model.",
    "dpdfnet4": "Higher quality DPDFNet-4 model.",
    "dpdfnet8": "Highest quality 16 kHz DPDFNet-8 model.",
    "dpdfnet2_48khz_hr": "High-resolution 48 kHz DPDFNet-2 model.",
    "dpdfnet8_48khz_hr": "High-resolution 48 kHz DPDFNet-8 model.",
}
MODEL_REGISTRY: Dict[str, ModelInfo] = {
    n: ModelInfo(name=n, sample_rate=s.sample_rate, frame_ms=20.0, description=_DESCRIPTIONS[n], onnx_filename=f"{n}.onnx")
    for n, s in MODEL_SPECS.items()
}
DEFAULT_MODEL = "dpdfnet2"
RANDOM_WEIGHTS = Path("<seeded-random-weights>")


@dataclass(frozen=True)
class ResolvedModel:
    info: ModelInfo
    onnx_path: Path          # reference field name; holds the .pth / .dpdfw path (or RANDOM_WEIGHTS)


def supported_models() -> List[str]:
    return list(MODEL_REGISTRY)


def get_model_info(model: str) -> ModelInfo:
    try:
        return MODEL_REGISTRY[model]
    except KeyError:
        raise ValueError(f"Unknown model '{model}'. Supported models: {', '.join(MODEL_REGISTRY)}") from None


def _candidates(name: str) -> List[Path]:
    root = os.environ.get("DPDFNET_MODEL_DIR")
    if not root:
        return []
    r = Path(root).expanduser()
    return [r / f"{name}.pth", r / "checkpoints" / f"{name}.pth", r / f"{name}.dpdfw"]


def available_model_entries() -> List[Dict[str, Any]]:
    rows = []
    for info in MODEL_REGISTRY.values():
        found = next((p for p in _candidates(info.name) if p.is_file()), None)
        row = asdict(info)
        row.update(weights_found=found is not None, weights_path=str(found) if found else None,
                   onnx_found=False, cached=found is not None)
        rows.append(row)
    return rows


def resolve_model(model: str = DEFAULT_MODEL, onnx_path: Optional[Union[str, Path]] = None,
                  auto_download: bool = True, verbose: bool = False) -> ResolvedModel:
    info = get_model_info(model)
    if onnx_path is not None:
        p = Path(onnx_path).expanduser()
        if not p.is_file():
            raise FileNotFoundError(f"Model weights file not found: {p}")
        return ResolvedModel(info=info, onnx_path=p)
    for p in _candidates(info.name):
        if p.is_file():
            if verbose:
                print(f"[dpdfnet-b200] using {p}")
            return ResolvedModel(info=info, onnx_path=p)
    if os.environ.get("DPDFNET_B200_RANDOM_WEIGHTS") == "1":
        return ResolvedModel(info=info, onnx_path=RANDOM_WEIGHTS / info.name)
    raise FileNotFoundError(
        f"No weights for model '{model}'. Put the reference checkpoint at $DPDFNET_MODEL_DIR/{info.name}.pth "
        "(this engine does not download; there is no network path in scope).")
