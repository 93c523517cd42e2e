"""API-contract tests for the drop-in layer, run on CPU with fake runtime sessions.

They restate the behaviours the reference pins in ``package/tests/test_package_behaviors.py``
(:341-403 buffering / flush / reset, :474-520 passthrough reconstruction and block-size invariance,
:523-609 offline alignment, :612-634 sample-rate change / empty chunk / stereo, :73-179 progress
callback and attenuation limit, :641-794 audio helpers) against this package, using the same
monkeypatch seams (``stream.resolve_model``, ``stream.build_runtime_model``, ``stream.infer_win_len``).
"""
from pathlib import Path

import numpy as np
import pytest

import dpdfnet_b200
from dpdfnet_b200 import api, audio, models, onnx_backend, stream


class _IO:
    def __init__(self, name, shape=None):
        self.name, self.shape = name, shape


class _Session:
    def __init__(self, bins, passthrough):
        self.bins, self.passthrough, self.calls = bins, passthrough, 0

    def get_inputs(self):
        return [_IO("spec", (1, 1, self.bins, 2)), _IO("state")]

    def get_outputs(self):
        return [_IO("out_spec"), _IO("out_state")]

    def run(self, _names, feed):
        self.calls += 1
        if self.passthrough:
            return feed["spec"].copy(), feed["state"].copy()
        return np.zeros((1, 1, self.bins, 2), np.float32), np.zeros(1, np.float32)


def _runtime(win, passthrough):
    return onnx_backend.RuntimeModel(session=_Session(win // 2 + 1, passthrough), init_state=np.zeros(1, np.float32),
                                     in_spec_name="spec", in_state_name="state", out_spec_name="out_spec",
                                     out_state_name="out_state")


def _enhancer(monkeypatch, win=320, passthrough=False, sr=16000):
    info = models.ModelInfo("dpdfnet2", sr, 20.0, "", "x.onnx")
    monkeypatch.setattr(stream, "resolve_model", lambda **kw: models.ResolvedModel(info, Path("fake.onnx")))
    monkeypatch.setattr(stream, "build_runtime_model", lambda p: _runtime(win, passthrough))
    monkeypatch.setattr(stream, "infer_win_len", lambda s, r: win)
    return dpdfnet_b200.StreamEnhancer(model="dpdfnet2")


def _stream_all(e, x, block):
    parts = [e.process(x[i:i + block], sample_rate=16000) for i in range(0, len(x), block)]
    parts.append(e.flush())
    return np.concatenate(parts)


def test_import_surface():
    for name in ("enhance", "available_models", "StreamEnhancer"):
        assert hasattr(dpdfnet_b200, name)
    dpdfnet_b200.install_as_dpdfnet()
    import dpdfnet
    from dpdfnet import stream as s2
    assert s2 is stream and dpdfnet.StreamEnhancer is stream.StreamEnhancer
    assert set(models.MODEL_REGISTRY) == {"baseline", "dpdfnet2", "dpdfnet4", "dpdfnet8", "dpdfnet2_48khz_hr", "dpdfnet8_48khz_hr"}
    assert models.MODEL_REGISTRY["dpdfnet8_48khz_hr"].sample_rate == 48000


def test_no_output_before_one_window_then_hop_per_frame(monkeypatch):
    e = _enhancer(monkeypatch, win=8)
    assert e.process(np.zeros(3, np.float32), sample_rate=16000).size == 0
    assert e.process(np.zeros(5, np.float32), sample_rate=16000).size == 4


def test_misaligned_chunks_neither_drop_nor_duplicate(monkeypatch):
    e = _enhancer(monkeypatch, win=320)
    total, fed, got = 16000, 0, 0
    while fed < total:
        n = min(171, total - fed)
        out = e.process(np.zeros(n, np.float32), sample_rate=16000)
        assert out.dtype == np.float32
        got += out.size
        fed += n
    assert got == ((total - 320) // 160 + 1) * 160


def test_reset_flush_empty_and_stereo(monkeypatch):
    e = _enhancer(monkeypatch, win=8)
    e.process(np.zeros(5, np.float32), sample_rate=16000)
    e.reset()
    assert e.process(np.zeros(5, np.float32), sample_rate=16000).size == 0
    out = e.flush()
    assert out.size > 0 and out.dtype == np.float32
    fresh = _enhancer(monkeypatch, win=8)
    assert fresh.flush().size == 0 and fresh.flush().dtype == np.float32
    assert fresh.process(np.zeros(0, np.float32), sample_rate=16000).size == 0
    assert fresh.process(np.zeros((8, 2), np.float32), sample_rate=16000).ndim == 1


def test_sample_rate_change_raises(monkeypatch):
    e = _enhancer(monkeypatch, win=8)
    e.process(np.zeros(4, np.float32), sample_rate=16000)
    with pytest.raises(ValueError, match="Sample rate changed"):
        e.process(np.zeros(4, np.float32), sample_rate=8000)
    e.reset()
    e.process(np.zeros(4, np.float32), sample_rate=8000)


def test_passthrough_reconstructs_input(monkeypatch):
    e = _enhancer(monkeypatch, win=320, passthrough=True)
    x = (np.random.default_rng(123).standard_normal(8000) * 0.5).astype(np.float32)
    y = e.process(x, sample_rate=16000)
    np.testing.assert_allclose(y[160:], x[160:y.size], atol=1e-5)


@pytest.mark.parametrize("block", [7, 64, 160, 171, 320, 512, 1000])
def test_block_size_invariance(monkeypatch, block):
    x = (np.random.default_rng(42).standard_normal(4000) * 0.5).astype(np.float32)
    ref = _stream_all(_enhancer(monkeypatch, passthrough=True), x, 1)
    got = _stream_all(_enhancer(monkeypatch, passthrough=True), x, block)
    assert got.size == ref.size
    np.testing.assert_allclose(got, ref, atol=1e-5)


def test_offline_passthrough_alignment_and_progress(monkeypatch):
    win, sr = 320, 16000
    info = models.ModelInfo("dpdfnet2", sr, 20.0, "", "x.onnx")
    rt = _runtime(win, True)
    monkeypatch.setattr(api, "resolve_model", lambda **kw: models.ResolvedModel(info, Path("fake.onnx")))
    monkeypatch.setattr(onnx_backend, "build_runtime_model", lambda p: rt)
    monkeypatch.setattr(onnx_backend, "infer_win_len", lambda s, r: win)
    x = (np.random.default_rng(7).standard_normal(6400) * 0.3).astype(np.float32)
    calls = []
    y = dpdfnet_b200.enhance(x, sr, progress_callback=lambda d, t: calls.append((d, t)))
    assert y.shape == x.shape and y.dtype == np.float32
    # the offline path is advanced by 2*win relative to its input (audio.py:133-136)
    np.testing.assert_allclose(y[:x.size - 2 * win], x[2 * win:], atol=1e-4)
    T = calls[0][1]
    assert calls[0] == (0, T) and calls[-1] == (T, T) and len(calls) == T + 1 and rt.session.calls == T


def test_attn_limit_blends_delayed_noisy_spectrum():
    rng = np.random.default_rng(0)
    noisy = rng.standard_normal((1, 9, 5, 2)).astype(np.float32)
    enh = rng.standard_normal((1, 9, 5, 2)).astype(np.float32)
    assert audio.apply_attn_limit(noisy, enh, None) is not None
    np.testing.assert_array_equal(audio.apply_attn_limit(noisy, enh, None), enh)
    zero_db = audio.apply_attn_limit(noisy, enh, 0.0)
    np.testing.assert_allclose(zero_db[:, 4:], noisy[:, :-4], atol=1e-7)
    assert not zero_db[:, :4].any()
    a = 10 ** (-12 / 20)
    mid = audio.apply_attn_limit(noisy, enh, 12.0)
    np.testing.assert_allclose(mid[:, 4:], a * noisy[:, :-4] + (1 - a) * enh[:, 4:], atol=1e-6)
    np.testing.assert_allclose(audio.apply_attn_limit(noisy, enh, float("inf")), enh, atol=1e-7)
    for bad in (-1.0, float("nan")):
        with pytest.raises(ValueError):
            audio.apply_attn_limit(noisy, enh, bad)


def test_audio_helpers():
    assert audio.to_mono(np.ones(4)).dtype == np.float32
    np.testing.assert_allclose(audio.to_mono(np.array([[1.0, 3.0], [2.0, 4.0]])), [2.0, 3.0])
    with pytest.raises(ValueError):
        audio.to_mono(np.zeros((2, 2, 2)))
    np.testing.assert_array_equal(audio.fit_length(np.arange(5.0), 3), [0, 1, 2])
    np.testing.assert_array_equal(audio.fit_length(np.arange(2.0), 4), [0, 1, 0, 0])
    np.testing.assert_array_equal(audio.pcm16_safe(np.array([2.0, -2.0, 0.5])), [32767, -32767, 16383])
    w = audio.vorbis_window(320)
    assert w.dtype == np.float32 and w.shape == (320,)
    np.testing.assert_allclose(w[:160] ** 2 + w[160:] ** 2, 1.0, atol=1e-6)
    cfg = audio.make_stft_config(960)
    assert cfg.hop_size == 480 and cfg.window.shape == (960,)
    x = np.arange(8, dtype=np.float32)
    assert audio.ensure_sample_rate(x, 16000, 16000) is not None and audio.ensure_sample_rate(x, 16000, 16000).size == 8
    assert audio.ensure_sample_rate(np.zeros(480, np.float32), 48000, 16000).size == 160


def test_model_resolution(monkeypatch, tmp_path):
    with pytest.raises(ValueError):
        models.get_model_info("nope")
    monkeypatch.delenv("DPDFNET_MODEL_DIR", raising=False)
    monkeypatch.delenv("DPDFNET_B200_RANDOM_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        models.resolve_model("dpdfnet2")
    monkeypatch.setenv("DPDFNET_MODEL_DIR", str(tmp_path))
    (tmp_path / "dpdfnet4.pth").write_bytes(b"x")
    assert models.resolve_model("dpdfnet4").onnx_path == tmp_path / "dpdfnet4.pth"
    rows = {r["name"]: r for r in models.available_model_entries()}
    assert rows["dpdfnet4"]["weights_found"] and not rows["dpdfnet2"]["weights_found"]
    with pytest.raises(FileNotFoundError):
        models.resolve_model("dpdfnet2", onnx_path=tmp_path / "missing.pth")
    monkeypatch.setenv("DPDFNET_B200_RANDOM_WEIGHTS", "1")
    assert models.resolve_model("dpdfnet2").onnx_path.parent == models.RANDOM_WEIGHTS


def test_initial_state_layout():
    from dpdfnet_b200.spec import get_spec
    st = onnx_backend.initial_state(get_spec("dpdfnet2"))
    assert st.shape == (45424,) and st[0] == -60.0 and abs(st[31] + 90.0) < 1e-4
    assert abs(st[32] - 1e-3) < 1e-9 and not st[128:].any()
    assert onnx_backend.infer_win_len(_Session(481, False), 48000) == 960
