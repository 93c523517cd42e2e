"""Produce a REAL ``.onnx`` export of the reference streaming graph offline (test infrastructure only).

The reference's exporter (onnx_model/export_dpdfnet_to_onnx.py:114-175) calls ``torch.onnx.export`` on
``DPDFNetOnnxWrapper(model)`` and then attaches metadata with the ``onnx`` package.  ``onnx`` is not in the image, but
torch's TorchScript exporter serialises the ModelProto in C++ and only needs ``onnx`` for a no-op post-processing hook
(custom onnxscript functions: there are none), so this module runs the same export with that hook bypassed and appends
the reference's own ``build_meta_data(model)`` entries as raw protobuf ``metadata_props`` fields (serialised messages
concatenate).  The result is what ``dpdfnet_b200/onnx_ingest.py`` must be able to read: folded BatchNorms, ONNX ``GRU``
nodes, anonymous MatMul constants and all.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import sys
import warnings
from pathlib import Path


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def metadata_props(meta: dict) -> bytes:
    """ModelProto.metadata_props (field 14) entries: StringStringEntryProto{key=1, value=2}."""
    return b"".join(_ld(14, _ld(1, str(k).encode()) + _ld(2, str(v).encode())) for k, v in meta.items())


def export_reference_onnx(spec, checkpoint, path, opset: int = 17, with_metadata: bool = True) -> Path:
    import torch
    from oracle import ref_import
    path = Path(path)
    wrapper, model = ref_import.export_wrapper(spec, checkpoint)
    exp = importlib.import_module("onnx_model.export_dpdfnet_48khz_hr_to_onnx" if spec.hr48 else "onnx_model.export_dpdfnet_to_onnx")
    stub = sys.modules.pop("onnx", None)                    # the placeholder ref_import installed must not look like the real package
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    saved = onnx_proto_utils._add_onnxscript_fn
    onnx_proto_utils._add_onnxscript_fn = lambda proto, opsets: proto
    try:
        x = torch.randn(1, 1, model.freq_bins, 2, dtype=torch.float32)
        s = model.initial_state(dtype=torch.float32)
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            warnings.simplefilter("ignore")
            torch.onnx.export(wrapper, (x, s), f=str(path), input_names=["spec", "state_in"], output_names=["spec_e", "state_out"],
                              opset_version=opset, do_constant_folding=True, dynamo=False)      # export...:118-137 (legacy branch)
    finally:
        onnx_proto_utils._add_onnxscript_fn = saved
        if stub is not None:
            sys.modules["onnx"] = stub
    if with_metadata:
        with open(path, "ab") as f:
            f.write(metadata_props(exp.build_meta_data(model)))                                  # export...:59-83
    return path
