"""Import helpers for the reference PyTorch modules (only where /root/reference is mounted).

TEST INFRASTRUCTURE ONLY: used by oracle/make_golden.py and the oracle-vs-reference tests that
are skipped when the reference tree is absent (e.g. on the GPU box).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF = os.environ.get("DPDFNET_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "onnx_model"))


def _prep():
    if "soundfile" not in sys.modules:          # onnx_model/dpdfnet.py:1 imports it at top level
        try:
            import soundfile  # noqa: F401
        except Exception:
            sys.modules["soundfile"] = types.ModuleType("soundfile")
    for p in (REF, os.path.join(REF, "model")):
        if p not in sys.path:
            sys.path.insert(0, p)


def streaming_model(spec, checkpoint):
    """Reference per-frame model (onnx_model/) loaded with ``checkpoint`` (offline naming)."""
    import torch
    _prep()
    with contextlib.redirect_stdout(io.StringIO()):
        if spec.hr48:
            from onnx_model.dpdfnet_48khz_hr import DPDFNet48HR as M, correct_state_dict
        else:
            from onnx_model.dpdfnet import DPDFNet as M, correct_state_dict
        m = M(dprnn_num_blocks=spec.n_blocks)
    sd = {k: torch.from_numpy(v.copy()) for k, v in checkpoint.items()}
    missing, unexpected = m.load_state_dict(correct_state_dict(sd), strict=False)
    learned_missing = [k for k in missing if not any(s in k for s in ("erb_fb", "erb_inv_fb", "stft.", "istft", "num_batches"))]
    assert not learned_missing and not unexpected, (learned_missing, unexpected)
    return m.eval()


def offline_model(spec, checkpoint):
    """Reference whole-utterance model (model/) loaded with ``checkpoint``."""
    import torch
    _prep()
    import importlib.util
    # load by path under a private name: a bare ``import dpdfnet`` may already be taken by the drop-in alias
    fname, cls = ("dpdfnet_48khz_hr.py", "DPDFNet48HR") if spec.hr48 else ("dpdfnet.py", "DPDFNet")
    modname = "_ref_offline_" + fname[:-3]
    if modname not in sys.modules:
        sp = importlib.util.spec_from_file_location(modname, os.path.join(REF, "model", fname))
        mod = importlib.util.module_from_spec(sp)
        sys.modules[modname] = mod
        sp.loader.exec_module(mod)
    M = getattr(sys.modules[modname], cls)
    with contextlib.redirect_stdout(io.StringIO()):
        m = M(dprnn_num_blocks=spec.n_blocks)
    sd = {k: torch.from_numpy(v.copy()) for k, v in checkpoint.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    learned_missing = [k for k in missing if not any(s in k for s in ("erb_fb", "erb_inv_fb", "stft.", "istft", "num_batches"))]
    assert not learned_missing and not unexpected, (learned_missing, unexpected)
    return m.eval()
