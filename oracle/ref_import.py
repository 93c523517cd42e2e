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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_ref() -> str:
    """$DPDFNET_REFERENCE, else the mounted tree, else the verbatim copy made by oracle/build_ref.py."""
    env = os.environ.get("DPDFNET_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/onnx_model"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF = _resolve_ref()


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "onnx_model"))


def kind() -> str:
    return "mounted reference tree" if not REF.startswith(_HERE) else "oracle/_ref (verbatim copy, oracle/build_ref.py)"


def _stub(name: str):
    """Placeholder for a third-party module the reference imports at top level but that is not in this image
    (no network).  Only modules whose *arithmetic is not on the path being run* are stubbed."""
    if name in sys.modules:
        return
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = types.ModuleType(name)


def _prep():
    _stub("soundfile")          # onnx_model/dpdfnet.py:1 imports it at top level (file I/O only)
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


# ---------------------------------------------------------------------------------------------
# The reference's PUBLIC streaming API (package/src/dpdfnet/stream.py:StreamEnhancer) on the reference's own
# per-frame torch module.  onnxruntime and the .onnx files are absent offline, so the ORT session behind
# RuntimeModel is replaced by the module tree the exporter traces (DPDFNetOnnxWrapper around onnx_model DPDFNet,
# export_dpdfnet_to_onnx.py:14-25, grouped linears converted to einsum form as the exporter does, :108) executed
# eagerly by torch on one thread - the fall-back SURVEY.md section 8(d) names.  Everything else (framing, rfft,
# session.run call shape, irfft, overlap-add) is the reference's unmodified code.
# ---------------------------------------------------------------------------------------------
class _NodeArg:
    def __init__(self, name, shape):
        self.name, self.shape = name, shape


class TorchSession:
    """``onnxruntime.InferenceSession`` look-alike around the reference's export wrapper module."""

    def __init__(self, wrapper, freq_bins: int, state_size: int):
        self._m = wrapper
        self._in = [_NodeArg("spec", [1, 1, freq_bins, 2]), _NodeArg("state_in", [state_size])]
        self._out = [_NodeArg("spec_e", [1, 1, freq_bins, 2]), _NodeArg("state_out", [state_size])]

    def get_inputs(self):
        return self._in

    def get_outputs(self):
        return self._out

    def run(self, output_names, feed):
        import torch
        with torch.no_grad():
            y, s = self._m(torch.from_numpy(feed["spec"]), torch.from_numpy(feed["state_in"]))
        return [y.numpy(), s.numpy()]


def export_wrapper(spec, checkpoint):
    """The module the reference exports to ONNX: wnorm scaling around the streaming model, einsum grouped linears."""
    import importlib
    _prep()
    _stub("onnx")               # export_dpdfnet_to_onnx.py:6 (serialisation only)
    m = streaming_model(spec, checkpoint)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module("onnx_model.export_dpdfnet_48khz_hr_to_onnx" if spec.hr48 else "onnx_model.export_dpdfnet_to_onnx")
        layers = importlib.import_module("onnx_model.layers")
    layers.convert_grouped_linear_to_einsum(m)
    wrapper_cls = [getattr(mod, n) for n in dir(mod) if n.endswith("OnnxWrapper")][0]
    return wrapper_cls(m).eval(), m


def reference_package():
    """The reference's ``dpdfnet`` package loaded from its source tree under the private name ``_ref_dpdfnet``
    (the name ``dpdfnet`` may be taken by this repo's drop-in alias)."""
    import importlib.util
    if "_ref_dpdfnet" in sys.modules:
        return sys.modules["_ref_dpdfnet"]
    _stub("librosa")            # audio.py:5 - resampling / offline STFT helpers, not used by the causal stream path at model SR
    _stub("onnxruntime")        # onnx_backend.py:8 - the session is supplied by TorchSession
    pkg_dir = os.path.join(REF, "package", "src", "dpdfnet")
    sp = importlib.util.spec_from_file_location("_ref_dpdfnet", os.path.join(pkg_dir, "__init__.py"),
                                                submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(sp)
    sys.modules["_ref_dpdfnet"] = mod
    sp.loader.exec_module(mod)
    return mod


def reference_stream_enhancer(spec, checkpoint):
    """An unmodified reference ``StreamEnhancer`` whose runtime is the reference torch per-frame graph.  Uses the
    monkeypatch seams the reference's own tests use (package/tests/test_package_behaviors.py:291-323)."""
    import importlib
    import numpy as np
    import torch
    reference_package()
    stream = importlib.import_module("_ref_dpdfnet.stream")
    backend = importlib.import_module("_ref_dpdfnet.onnx_backend")
    wrapper, model = export_wrapper(spec, checkpoint)
    sess = TorchSession(wrapper, spec.freq_bins, model.state_size())
    init = model.initial_state(dtype=torch.float32).numpy().copy()
    runtime = backend.RuntimeModel(session=sess, init_state=np.ascontiguousarray(init), in_spec_name="spec",
                                   in_state_name="state_in", out_spec_name="spec_e", out_state_name="state_out")

    class _Info:
        sample_rate = spec.sample_rate

    class _Resolved:
        onnx_path = "<torch reference graph>"
        info = _Info()

    saved = stream.resolve_model, stream.build_runtime_model
    stream.resolve_model = lambda **kw: _Resolved()
    stream.build_runtime_model = lambda path: runtime
    try:
        se = stream.StreamEnhancer(model=spec.name)
    finally:
        stream.resolve_model, stream.build_runtime_model = saved
    return se
