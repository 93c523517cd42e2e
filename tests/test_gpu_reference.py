"""CUDA engine against the LIVE reference code (not the oracle): the reference's public StreamEnhancer
(package/src/dpdfnet/stream.py) driving the reference's per-frame torch graph, from /root/reference or from the verbatim
copy oracle/_ref that travels to the GPU box (oracle/build_ref.py).  Tolerance: north_star's 1e-4 on the waveform."""
import numpy as np
import pytest

from dpdfnet_b200.spec import get_spec
from dpdfnet_b200.weights import random_checkpoint

pytestmark = [pytest.mark.gpu, pytest.mark.reference]
WAVE_TOL = 1e-4


@pytest.mark.parametrize("name,arm", [("dpdfnet2", "ffma2"), ("dpdfnet4", "tcgen05"), ("dpdfnet2_48khz_hr", "tcgen05")])
def test_engine_stream_matches_reference_stream_enhancer(name, arm):
    import torch
    from dpdfnet_b200.engine import Engine
    from oracle import ref_import
    torch.set_num_threads(max(1, torch.get_num_threads()))
    spec = get_spec(name)
    ck = random_checkpoint(spec, 0)
    se = ref_import.reference_stream_enhancer(spec, ck)
    hops = 30
    rng = np.random.default_rng(21)
    x = np.clip(rng.standard_normal((hops + 1) * spec.hop) * 0.1, -1, 1).astype(np.float32)
    ref = se.process(x)                                            # the reference's own causal STFT -> model -> iSTFT/OLA
    assert ref.size == hops * spec.hop
    B = 3
    eng = Engine(spec, ck, max_streams=B)
    if arm == "tcgen05":
        for k in ("intra_tc", "post_tc", "sep_tc", "gru_tc", "dft_tc"):
            eng.set_option(k, 1)
    pcm = np.tile(x[None], (B, 1))
    pcm[1] *= 0.5                                                  # the neighbours carry different signals
    pcm[2] = pcm[2][::-1]
    eng.prime_pcm_host(pcm[:, :spec.hop])
    out = eng.run_pcm_host(pcm[:, spec.hop:])
    err = float(np.abs(out[0] - ref).max())
    assert err < WAVE_TOL, err
    # flat state in the reference layout against the reference session's state vector
    assert np.abs(eng.state_export(0) - se._state).max() < 5e-5
    eng.close()


def test_stream_enhancer_runs_a_real_onnx_export(tmp_path):
    """f3 end to end: reference graph -> torch ONNX exporter -> .onnx file -> StreamEnhancer(onnx_path=...) on the engine
    (weights from the initialisers, initial state from the metadata) == the reference StreamEnhancer on the same weights."""
    import dpdfnet_b200
    from dpdfnet_b200.onnx_backend import EnginePool
    from oracle import ref_import
    from oracle.onnx_export import export_reference_onnx
    spec = get_spec("dpdfnet2")
    ck = random_checkpoint(spec, 4)
    path = export_reference_onnx(spec, ck, tmp_path / "dpdfnet2.onnx")
    x = np.clip(np.random.default_rng(8).standard_normal(21 * spec.hop) * 0.1, -1, 1).astype(np.float32)
    ref = ref_import.reference_stream_enhancer(spec, ck).process(x)
    e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2", onnx_path=path)
    assert e._fused
    got = e.process(x, 16000)
    assert got.shape == ref.shape and np.abs(got - ref).max() < WAVE_TOL
    # the ONNX-shaped seam with the metadata-built initial state
    rt = e._runtime
    y, st = rt.session.run([rt.out_spec_name, rt.out_state_name],
                           {rt.in_spec_name: np.zeros((1, 1, 161, 2), np.float32), rt.in_state_name: rt.init_state.copy()})
    assert st.shape == (spec.state_size,) and np.isfinite(y).all()
    e.close()
    EnginePool.shutdown()
