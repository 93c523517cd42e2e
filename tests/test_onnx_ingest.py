"""f3: reading the reference's shipped artefact, an ONNX export, without onnx / onnxruntime.
Error behaviour mirrors package/src/dpdfnet/onnx_backend.py:23-24, 52-78; the round trip runs on a REAL export of the
reference graph produced by torch's exporter (oracle/onnx_export.py) where the reference sources are available."""
import numpy as np
import pytest

from dpdfnet_b200 import onnx_ingest
from dpdfnet_b200.spec import get_spec
from dpdfnet_b200.weights import pack_tensors, random_checkpoint
from oracle.onnx_export import _ld, _varint, metadata_props


def _tensor_proto(name, arr):
    arr = np.asarray(arr, np.float32)
    dims = b"".join(_varint((1 << 3) | 0) + _varint(d) for d in arr.shape)
    return dims + _varint((2 << 3) | 0) + _varint(1) + _ld(8, name.encode()) + _ld(9, arr.tobytes())


def _value_info(name, shape):
    dims = b"".join(_ld(1, _varint((1 << 3) | 0) + _varint(d)) for d in shape)
    ttype = _varint((1 << 3) | 0) + _varint(1) + _ld(2, dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, ttype))


def _mini_model(inputs, meta, tensors=()):
    graph = b"".join(_ld(5, _tensor_proto(n, a)) for n, a in tensors) + b"".join(_ld(11, _value_info(n, s)) for n, s in inputs)
    return _varint((1 << 3) | 0) + _varint(8) + _ld(7, graph) + metadata_props(meta)


META16 = {"state_size": 45424, "erb_norm_state_size": 32, "spec_norm_state_size": 96, "sample_rate": 16000, "freq_bins": 161,
          "erb_norm_init": ",".join(["-60"] * 32), "spec_norm_init": ",".join(["0.001"] * 96)}


def test_wire_format_reader_and_metadata_state(tmp_path):
    p = tmp_path / "m.onnx"
    w = np.arange(12, dtype=np.float32).reshape(3, 4)
    p.write_bytes(_mini_model([("spec", (1, 1, 161, 2)), ("state_in", (45424,))], META16, [("model.some.weight", w)]))
    m = onnx_ingest.read_onnx(p)
    assert [i[0] for i in m.inputs] == ["spec", "state_in"] and m.inputs[0][1] == [1, 1, 161, 2]
    assert np.array_equal(m.initializers["model.some.weight"], w)
    st = onnx_ingest.initial_state_from_metadata(m)
    assert st.shape == (45424,) and st[0] == -60 and st[32] == np.float32(0.001) and not st[128:].any()
    spec = onnx_ingest.spec_from_metadata(m)
    assert spec.name == "dpdfnet2" and spec.n_blocks == 2
    assert onnx_ingest.spec_from_metadata(m, n_blocks=0).name == "baseline"
    with pytest.raises(ValueError, match="do not cover"):
        onnx_ingest.checkpoint_from_initializers(m, spec)


def test_errors_match_the_reference_backend(tmp_path):
    with pytest.raises(FileNotFoundError, match="ONNX model file not found"):
        onnx_ingest.read_onnx(tmp_path / "missing.onnx")
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff")
    with pytest.raises(ValueError, match="not a readable ONNX"):
        onnx_ingest.read_onnx(bad)
    one = tmp_path / "one.onnx"
    one.write_bytes(_mini_model([("spec", (1, 1, 161, 2))], META16))
    with pytest.raises(ValueError, match="two inputs"):
        onnx_ingest.initial_state_from_metadata(onnx_ingest.read_onnx(one))
    nometa = tmp_path / "nometa.onnx"
    meta = dict(META16)
    del meta["erb_norm_init"]
    nometa.write_bytes(_mini_model([("spec", (1, 1, 161, 2)), ("state_in", (45424,))], meta))
    with pytest.raises(ValueError, match="missing required metadata key: 'erb_norm_init'"):
        onnx_ingest.initial_state_from_metadata(onnx_ingest.read_onnx(nometa))
    odd = tmp_path / "odd.onnx"
    odd.write_bytes(_mini_model([("spec", (1, 1, 161, 2)), ("state_in", (45000,))], dict(META16, state_size=45000)))
    with pytest.raises(ValueError, match="does not match any DPDFNet configuration"):
        onnx_ingest.spec_from_metadata(onnx_ingest.read_onnx(odd))


@pytest.mark.reference
@pytest.mark.parametrize("name", ["dpdfnet2", "dpdfnet2_48khz_hr", "baseline"])
def test_real_export_round_trip(tmp_path, name):
    """Reference graph -> torch ONNX exporter -> file -> hand-rolled reader -> packed engine weights == the checkpoint's,
    and the metadata state == what the engine initialises slots with."""
    from dpdfnet_b200.onnx_backend import initial_state
    from oracle.onnx_export import export_reference_onnx
    spec = get_spec(name)
    ck = random_checkpoint(spec, 3)
    path = export_reference_onnx(spec, ck, tmp_path / f"{name}.onnx")
    m = onnx_ingest.read_onnx(path)
    assert m.metadata["state_size"] == str(spec.state_size)
    got_spec, sd = onnx_ingest.load_onnx_checkpoint(path)
    assert (got_spec.sample_rate, got_spec.n_blocks, got_spec.hr48) == (spec.sample_rate, spec.n_blocks, spec.hr48)
    a, b = pack_tensors(spec, ck), pack_tensors(spec, sd)
    for k in a:
        if ".tc" in k or k.endswith("tc_pw") or k.endswith("tc_w"):
            continue          # FP16 hi/lo operand images bit-packed into float32 words: derived from the tensors checked here
        assert np.abs(a[k] - b[k]).max() <= 1e-6 * (np.abs(a[k]).max() + 1e-12), k
    assert np.abs(onnx_ingest.initial_state_from_metadata(m) - initial_state(spec)).max() < 1e-6
