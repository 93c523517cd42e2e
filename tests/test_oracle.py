"""Pins the CPU oracle: against golden vectors produced by the reference PyTorch modules
(oracle/make_golden.py) and - in the build container - against the reference itself."""
import numpy as np
import pytest

from dpdfnet_b200.spec import MODEL_SPECS, get_spec
from dpdfnet_b200.weights import pack_tensors, random_checkpoint, ref_param_shapes
from oracle.oracle_np import OracleEngine, offline_exact

STREAM_TOL = 1e-4      # un-normalised spectrum units (|X| up to ~30): ~4e-6 relative
WAVE_TOL = 1e-4        # north_star tolerance on the enhanced waveform (measured ~3e-7)


def _engine(name, seed, B):
    spec = get_spec(name)
    return OracleEngine(spec, pack_tensors(spec, random_checkpoint(spec, seed)), B)


ALL_MODELS = ["baseline", "dpdfnet2", "dpdfnet4", "dpdfnet8", "dpdfnet2_48khz_hr", "dpdfnet8_48khz_hr"]


@pytest.mark.parametrize("name", ALL_MODELS)
def test_stream_golden(golden_dir, name):
    g = np.load(golden_dir / f"stream_{name}.npz")
    eng = _engine(name, int(g["seed"]), 1)
    for t in range(g["spec_in"].shape[0]):
        y = eng.step_spec(g["spec_in"][t][None])
        assert np.abs(y[0] - g["spec_out"][t]).max() < STREAM_TOL, t
    st = eng.export_state(0)[::int(g["state_stride"]) if "state_stride" in g else 1]
    assert st.shape == g["state"].shape
    assert np.abs(st - g["state"]).max() < 5e-5


def test_offline_golden_cfg0_10s_clip(golden_dir):
    """BASELINE.json configs[0]: dpdfnet2 16 kHz, one 10 s noisy clip against model/dpdfnet.py (1e-4 on the waveform)."""
    from oracle.make_golden import test_signal
    g = np.load(golden_dir / "offline_cfg0_dpdfnet2_10s.npz")
    spec = get_spec("dpdfnet2")
    wave = test_signal(np.random.default_rng(int(g["signal_seed"])), spec.sample_rate, int(g["seconds"]) * spec.sample_rate, 1)
    out = offline_exact(_engine("dpdfnet2", int(g["seed"]), 1), wave)
    assert out.shape == g["wave_out"].shape == (1, 160000)
    assert np.abs(out - g["wave_out"]).max() < WAVE_TOL


@pytest.mark.parametrize("tag", ["fullscale", "quiet"])
def test_offline_golden_level_extremes(golden_dir, tag):
    """Clipped full-scale input (RMS 0.48) and a -100 dBFS input through the offline reference."""
    from oracle.make_golden import test_signal
    g = np.load(golden_dir / f"offline_dpdfnet2_{tag}.npz")
    spec = get_spec("dpdfnet2")
    wave = np.clip(test_signal(np.random.default_rng(int(g["signal_seed"])), spec.sample_rate, spec.sample_rate, 1) * float(g["gain"]), -1, 1).astype(np.float32)
    out = offline_exact(_engine("dpdfnet2", int(g["seed"]), 1), wave)
    assert np.abs(out - g["wave_out"]).max() < WAVE_TOL


@pytest.mark.parametrize("name", ALL_MODELS)
def test_offline_golden(golden_dir, name):
    g = np.load(golden_dir / f"offline_{name}.npz")
    eng = _engine(name, int(g["seed"]), g["wave_in"].shape[0])
    out = offline_exact(eng, g["wave_in"])
    assert out.shape == g["wave_out"].shape
    assert np.abs(out - g["wave_out"]).max() < WAVE_TOL


def test_state_roundtrip_and_slot_independence():
    eng = _engine("dpdfnet2", 3, 3)
    rng = np.random.default_rng(0)
    F = eng.spec.freq_bins
    X = (rng.standard_normal((5, 3, F, 2)) * 10).astype(np.float32)
    for t in range(3):
        eng.step_spec(X[t])
    flat = eng.export_state(1)
    assert flat.size == eng.spec.state_size
    other = _engine("dpdfnet2", 3, 1)
    other.import_state(0, flat)
    a = eng.step_spec(X[3])
    b = other.step_spec(X[3][1:2])
    assert np.abs(a[1] - b[0]).max() < 1e-4      # BLAS blocking differs with batch size
    # slot subset stepping leaves the other slots untouched
    before = eng.export_state(0)
    eng.step_spec(X[4][1:3], slots=np.array([1, 2]))
    assert np.array_equal(before, eng.export_state(0))


def test_state_sizes_match_reference_table():
    # SURVEY.md section 8a totals (verified against model.state_size())
    expect = {"dpdfnet2": 45424, "dpdfnet4": 52592, "dpdfnet8": 66928,
              "dpdfnet2_48khz_hr": 56436, "dpdfnet8_48khz_hr": 90228}
    for k, v in expect.items():
        assert get_spec(k).state_size == v


def test_param_counts_match_readme():
    # README.md:31-41 parameter counts (M); lsnr head and buffers included as in the reference count
    expect = {"dpdfnet2": 2.498, "dpdfnet4": 2.848, "dpdfnet8": 3.548,
              "dpdfnet2_48khz_hr": 2.583, "dpdfnet8_48khz_hr": 3.634}
    for k, v in expect.items():
        n = sum(int(np.prod(s)) for name, s in ref_param_shapes(get_spec(k)).items() if "running" not in name)
        assert abs(n / 1e6 - v) < 2e-3, (k, n)


def test_erb_bands():
    w16 = get_spec("dpdfnet2").erb_widths
    assert w16 == [1] * 11 + [2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 7, 7, 8, 8, 10, 12, 12, 14, 16, 18]
    w48 = get_spec("dpdfnet2_48khz_hr").erb_widths
    assert sum(w48) == 481 and w48[:13] == [2] * 13


@pytest.mark.reference
@pytest.mark.parametrize("name", ["dpdfnet2", "dpdfnet8_48khz_hr"])
def test_against_reference_streaming(name):
    import torch
    from oracle import ref_import
    spec = get_spec(name)
    ck = random_checkpoint(spec, 7)
    ref = ref_import.streaming_model(spec, ck)
    assert ref.state_size() == spec.state_size
    eng = OracleEngine(spec, pack_tensors(spec, ck), 1)
    rng = np.random.default_rng(5)
    state = ref.initial_state(dtype=torch.float32)
    assert np.array_equal(state.numpy(), eng.export_state(0))
    for t in range(6):
        X = (rng.standard_normal((1, spec.freq_bins, 2)) * 20).astype(np.float32)
        with torch.no_grad():
            out, state = ref(torch.from_numpy(X)[None] * np.float32(spec.wnorm), state)
            out = out * np.float32(1 / spec.wnorm)
        y = eng.step_spec(X)
        assert np.abs(out.numpy()[0, 0] - y[0]).max() < STREAM_TOL
    assert np.abs(state.numpy() - eng.export_state(0)).max() < 5e-5


@pytest.mark.reference
def test_against_reference_offline_cfg0():
    """BASELINE.json configs[0]: dpdfnet2 16 kHz, one clip, vs model/dpdfnet.py (shortened to 3 s)."""
    import torch
    from oracle import ref_import
    from oracle.make_golden import test_signal
    spec = get_spec("dpdfnet2")
    ck = random_checkpoint(spec, 0)
    off = ref_import.offline_model(spec, ck)
    wave = test_signal(np.random.default_rng(9), spec.sample_rate, 3 * spec.sample_rate, 1)
    with torch.no_grad():
        ref_out, _ = off(torch.from_numpy(wave))
    eng = OracleEngine(spec, pack_tensors(spec, ck), 1)
    out = offline_exact(eng, wave)
    assert np.abs(out - ref_out.numpy()).max() < WAVE_TOL
