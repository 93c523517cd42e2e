"""Drop-in API on the GPU: StreamEnhancer's fused path, the ONNX-shaped EngineSession and the batched
offline entry, each against the oracle / golden vectors (random seeded weights, seed 0)."""
import numpy as np
import pytest

from dpdfnet_b200.spec import get_spec
from dpdfnet_b200.weights import pack_tensors, random_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture()
def random_weights(monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    monkeypatch.setenv("DPDFNET_B200_RANDOM_WEIGHTS", "1")
    monkeypatch.delenv("DPDFNET_MODEL_DIR", raising=False)


def _oracle(name, B=1):
    from oracle.oracle_np import OracleEngine
    spec = get_spec(name)
    return OracleEngine(spec, pack_tensors(spec, random_checkpoint(spec, 0)), B)


def test_stream_enhancer_fused_matches_oracle_and_is_block_size_invariant(random_weights):
    import dpdfnet_b200
    x = (np.random.default_rng(3).standard_normal(4000) * 0.1).astype(np.float32)
    ora = _oracle("dpdfnet2")
    hop = 160
    ora.step_pcm(x[None, :hop], flags=np.array([4]))            # prime
    ref = np.concatenate([ora.step_pcm(x[None, t * hop:(t + 1) * hop])[0] for t in range(1, 4000 // hop)])
    outs = {}
    for block in (4000, 171, 7):
        e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
        assert e._fused
        parts = [e.process(x[i:i + block], sample_rate=16000) for i in range(0, x.size, block)]
        outs[block] = np.concatenate(parts)
        assert outs[block].size == ((4000 - 320) // 160 + 1) * 160
    assert np.abs(outs[4000] - ref).max() < 1e-4
    assert np.array_equal(outs[4000], outs[171]) and np.array_equal(outs[4000], outs[7])
    e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
    a = e.process(x[:1000], sample_rate=16000)
    tail = e.flush()
    assert tail.size == 160
    e.reset()
    b = e.process(x[:1000], sample_rate=16000)
    assert np.array_equal(a, b)


def test_engine_session_keeps_the_onnx_call_shape(random_weights, golden_dir):
    from dpdfnet_b200 import onnx_backend
    from dpdfnet_b200.models import resolve_model
    g = np.load(golden_dir / "stream_dpdfnet2.npz")
    rt = onnx_backend.build_runtime_model(resolve_model("dpdfnet2").onnx_path)
    assert onnx_backend.infer_win_len(rt.session, 16000) == 320
    assert [i.name for i in rt.session.get_inputs()] == [rt.in_spec_name, rt.in_state_name]
    state = rt.init_state.copy()
    for t in range(6):
        y, state = rt.session.run([rt.out_spec_name, rt.out_state_name],
                                  {rt.in_spec_name: g["spec_in"][t][None, None], rt.in_state_name: state})
        assert y.shape == (1, 1, 161, 2) and state.shape == (45424,)
        assert np.abs(y[0, 0] - g["spec_out"][t]).max() < 2e-4
    # a caller that copies the state (api.py:91 style) must get the same answer as one that does not
    y2, _ = rt.session.run([rt.out_spec_name, rt.out_state_name],
                           {rt.in_spec_name: g["spec_in"][6][None, None], rt.in_state_name: state.copy()})
    assert np.abs(y2[0, 0] - g["spec_out"][6]).max() < 2e-4
    with pytest.raises(ValueError):
        rt.session.run(None, {rt.in_spec_name: g["spec_in"][0][None, None], rt.in_state_name: state[:-1]})


def test_enhance_matches_host_pipeline_on_oracle(random_weights):
    """dpdfnet.enhance() on the engine == the same reference pipeline driven by the CPU oracle."""
    import dpdfnet_b200
    from dpdfnet_b200 import api, onnx_backend

    class OracleSession:
        def __init__(self):
            self.o = _oracle("dpdfnet2")

        def get_inputs(self):
            class I:
                shape = (1, 1, 161, 2)
            return [I()]

        def run(self, names, feed):
            self.o.import_state(0, feed["state_in"])
            y = self.o.step_spec(feed["spec"].reshape(1, 161, 2))
            return y.reshape(1, 1, 161, 2), self.o.export_state(0)

    x = (np.random.default_rng(5).standard_normal(8000) * 0.1).astype(np.float32)
    got = dpdfnet_b200.enhance(x, 16000, model="dpdfnet2", attn_limit_db=12.0)
    rt = onnx_backend.RuntimeModel(OracleSession(), onnx_backend.initial_state(get_spec("dpdfnet2")), "spec", "state_in", "spec_e", "state_out")
    ref = api._enhance_with_runtime(x, 16000, runtime=rt, model_sample_rate=16000, attn_limit_db=12.0)
    assert got.shape == x.shape
    assert np.abs(got - ref).max() < 1e-4


def test_enhance_batch_ragged_clips_match_offline_golden(random_weights, golden_dir):
    """Ragged clips in one batched run: every clip gets exactly what the offline model gives it alone (own reflect
    padding, own flush hops), the full-length one equals the reference PyTorch model's golden output end to end."""
    import dpdfnet_b200
    from dpdfnet_b200.offline import enhance_offline_exact
    from dpdfnet_b200.onnx_backend import create_session
    from dpdfnet_b200.models import resolve_model
    g = np.load(golden_dir / "offline_dpdfnet2.npz")
    clips = [g["wave_in"][0], g["wave_in"][1][:20000], g["wave_in"][1][:7531], g["wave_in"][0][:100]]
    out = dpdfnet_b200.enhance_batch(clips, 16000, model="dpdfnet2")
    assert [o.shape for o in out] == [c.shape for c in clips]
    assert np.abs(out[0] - g["wave_out"][0]).max() < 1e-4                    # whole clip, including its reflect-padded tail
    eng = create_session(resolve_model("dpdfnet2").onnx_path, max_streams=1).engine
    for i in (1, 2):
        alone = enhance_offline_exact(eng, clips[i][None])[0]
        assert np.abs(out[i][:alone.size] - alone).max() < 1e-6
        assert not out[i][alone.size:].any()                                 # the sub-hop remainder is not synthesised
    assert not out[3].any()                                                  # shorter than a hop: nothing to enhance
    lim = dpdfnet_b200.enhance_batch(clips[:2], 16000, model="dpdfnet2", attn_limit_db=12.0)
    alpha = 10.0 ** (-12.0 / 20.0)
    assert np.abs(lim[1] - (alpha * clips[1] + (1 - alpha) * out[1])).max() < 1e-6
    with pytest.raises(ValueError):
        dpdfnet_b200.enhance_batch(clips[:1], 16000, model="dpdfnet2", attn_limit_db=-1.0)


def test_enhance_batch_resamples_on_the_device(random_weights):
    """48 kHz clips through a 16 kHz model: device polyphase resampling in and out equals the host helper
    (scipy.signal.resample_poly) around the same engine run."""
    import dpdfnet_b200
    from dpdfnet_b200.audio import ensure_sample_rate
    rng = np.random.default_rng(5)
    clips = [(rng.standard_normal(n) * 0.1).astype(np.float32) for n in (24000, 15011)]
    got = dpdfnet_b200.enhance_batch(clips, 48000, model="dpdfnet2")
    assert [g.shape for g in got] == [c.shape for c in clips]
    down = [ensure_sample_rate(c, 48000, 16000) for c in clips]
    mid = dpdfnet_b200.enhance_batch(down, 16000, model="dpdfnet2")
    for g, m, c in zip(got, mid, clips):
        ref = ensure_sample_rate(m, 16000, 48000)
        n = min(ref.size, c.size)
        assert np.abs(g[:n] - ref[:n]).max() < 1e-4


def test_pipelined_host_entry_equals_the_synchronous_one(random_weights):
    """dpdf_submit_pcm_host / dpdf_wait (two hops in flight, copies on their own streams) == dpdf_step_pcm_host hop for hop,
    with slot indirection and per-row flags."""
    import torch
    from dpdfnet_b200.engine import Engine
    B, T = 300, 9
    eng_a, eng_b = Engine("dpdfnet2", None, max_streams=B + 4), Engine("dpdfnet2", None, max_streams=B + 4)
    hop = eng_a.spec.hop
    rng = np.random.default_rng(2)
    slots = rng.permutation(B + 4)[:B].astype(np.int32)
    pin_in = torch.from_numpy((rng.standard_normal((T, B, hop)) * 0.1).astype(np.float32)).pin_memory().numpy()
    pin_out = torch.empty(2, B, hop).pin_memory().numpy()
    flags = np.zeros(B, np.int32)
    flags[::7] = 1                                                # WARMUP rows on the first two hops
    ref = [eng_a.step_pcm_host(pin_in[t], slot_ids=slots, flags=flags if t < 2 else None).copy() for t in range(T)]
    got, prev = [], None
    for t in range(T):
        tk = eng_b.submit_pcm_host(pin_in[t], pin_out[t & 1], slot_ids=slots, flags=flags if t < 2 else None)
        if prev is not None:
            eng_b.wait(prev)
            got.append(pin_out[(t - 1) & 1].copy())
        prev = tk
    eng_b.wait(prev)
    got.append(pin_out[(T - 1) & 1].copy())
    for t in range(T):
        assert np.array_equal(got[t], ref[t]), t
    assert np.array_equal(eng_a.state_export(int(slots[5])), eng_b.state_export(int(slots[5])))
    with pytest.raises(ValueError):
        eng_b.wait(10 ** 6)
    eng_a.close()
    eng_b.close()
