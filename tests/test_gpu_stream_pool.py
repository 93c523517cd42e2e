"""Shared-engine streaming front-end on the GPU ("many StreamEnhancer objects may share one engine
via slot ids"; reference surface package/src/dpdfnet/stream.py:35-200): slot pooling, process_many, StreamGroup and the
stateful device resampler inside StreamEnhancer.process / flush."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def random_weights(monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    monkeypatch.setenv("DPDFNET_B200_RANDOM_WEIGHTS", "1")
    monkeypatch.setenv("DPDFNET_B200_POOL_STREAMS", "8")
    monkeypatch.delenv("DPDFNET_MODEL_DIR", raising=False)
    from dpdfnet_b200.onnx_backend import EnginePool
    EnginePool.shutdown()
    yield
    EnginePool.shutdown()


def _noise(n, seed, scale=0.1):
    return (np.random.default_rng(seed).standard_normal(n) * scale).astype(np.float32)


def test_enhancers_share_one_engine_and_slots_are_recycled(random_weights):
    import dpdfnet_b200
    from dpdfnet_b200.onnx_backend import EnginePool
    es = [dpdfnet_b200.StreamEnhancer(model="dpdfnet2") for _ in range(3)]
    engines = {id(e._runtime.session.engine) for e in es}
    assert len(engines) == 1 and len({e._runtime.session.slot for e in es}) == 3
    pool = next(iter(EnginePool._pools.values()))
    assert pool.capacity == 8 and pool.in_use == 3
    x = [_noise(1600, s) for s in range(3)]
    solo = [e.process(x[i], 16000) for i, e in enumerate(es)]
    # a fourth stream in a recycled slot starts from a clean state: same output as stream 0 got for the same input
    slot0 = es[0]._runtime.session.slot
    es[0].close()
    assert pool.in_use == 2
    again = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
    assert again._runtime.session.slot == slot0
    assert np.array_equal(again.process(x[0], 16000), solo[0])
    # the pool grows by a second engine when the first is full
    more = [dpdfnet_b200.StreamEnhancer(model="dpdfnet2") for _ in range(8)]
    assert len(pool.engines) == 2 and pool.in_use == 11
    for e in es[1:] + [again] + more:
        e.close()
    assert pool.in_use == 0


def test_process_many_equals_one_by_one(random_weights):
    """Ragged chunk sizes, streams joining late, flush: the batched call returns what the single calls return."""
    import dpdfnet_b200
    from dpdfnet_b200.stream import flush_many, process_many
    n = 5
    x = [_noise(4000 + 37 * i, 10 + i) for i in range(n)]
    sizes = [160, 171, 7, 1000, 320]
    ref = []
    for i in range(n):
        e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
        parts = [e.process(x[i][k:k + sizes[i]], 16000) for k in range(0, x[i].size, sizes[i])]
        parts.append(e.flush())
        ref.append(np.concatenate(parts))
        e.close()
    es = [dpdfnet_b200.StreamEnhancer(model="dpdfnet2") for _ in range(n)]
    got = [[] for _ in range(n)]
    cur = [0] * n
    while any(cur[i] < x[i].size for i in range(n)):
        live = [i for i in range(n) if cur[i] < x[i].size]
        outs = process_many([es[i] for i in live], [x[i][cur[i]:cur[i] + sizes[i]] for i in live], 16000)
        for i, o in zip(live, outs):
            got[i].append(o)
            cur[i] += sizes[i]
    for i, o in enumerate(flush_many(es)):
        got[i].append(o)
    for i in range(n):
        g = np.concatenate(got[i])
        assert g.shape == ref[i].shape
        assert np.abs(g - ref[i]).max() < 1e-6
    with pytest.raises(ValueError):
        process_many([es[0], es[0]], [x[0][:10], x[0][:10]], 16000)
    with pytest.raises(ValueError):                                   # sample-rate change without reset (stream.py:104-110)
        process_many(es[:1], [x[0][:10]], 48000)


def test_stream_group_rows_behave_like_stream_enhancers(random_weights):
    import dpdfnet_b200
    from dpdfnet_b200.stream import StreamGroup
    B, n = 6, 3000
    x = np.stack([_noise(n, 40 + b) for b in range(B)])
    ref = []
    for b in range(B):
        e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
        ref.append(np.concatenate([e.process(x[b], 16000), e.flush()]))
        e.close()
    ref = np.stack(ref)
    for block in (n, 160, 171, 7):
        g = StreamGroup(model="dpdfnet2", streams=B)
        parts = [g.process(x[:, k:k + block], 16000) for k in range(0, n, block)]
        assert parts[0].shape[1] == (0 if block < 320 else ((block - 320) // 160 + 1) * 160)      # nothing before one window
        parts.append(g.flush())
        got = np.concatenate(parts, 1)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-6, block
        g.reset()
        again = g.process(x[:, :1000], 16000)
        assert np.abs(again - ref[:, :again.shape[1]]).max() < 1e-6
        g.close()
    import torch
    g = StreamGroup(model="dpdfnet2", streams=B)
    y = g.process(torch.from_numpy(x).cuda(), 16000)                   # device tensors in -> device tensors out
    assert y.is_cuda and np.abs(y.cpu().numpy() - ref[:, :y.shape[1]]).max() < 1e-6
    with pytest.raises(ValueError):
        g.process(x[:, :10], 8000)
    g.close()


@pytest.mark.parametrize("sr", [48000, 8000])
def test_stateful_resampling_is_block_size_invariant(random_weights, sr):
    """f1: with a caller rate != model rate the engine path resamples statefully on the device, so the chunked result
    equals the one-shot result (the reference's stateless per-chunk librosa call, stream.py:112, cannot give that),
    and equals host resample_poly -> engine -> host resample_poly of the whole signal."""
    import dpdfnet_b200
    from dpdfnet_b200.audio import ensure_sample_rate
    from dpdfnet_b200.stream import StreamGroup
    n = sr // 2
    x = _noise(n, 77)
    one = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
    whole = np.concatenate([one.process(x, sr), one.flush()])
    one.close()
    for block in (sr // 100, 997):
        e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
        parts = [e.process(x[k:k + block], sr) for k in range(0, n, block)]
        parts.append(e.flush())
        got = np.concatenate(parts)
        e.close()
        assert got.shape == whole.shape and np.abs(got - whole).max() < 1e-6
    # against the host helper around a model-rate stream
    down = ensure_sample_rate(x, sr, 16000)
    e = dpdfnet_b200.StreamEnhancer(model="dpdfnet2")
    mid = np.concatenate([e.process(down, 16000), e.flush()])
    e.close()
    ref = ensure_sample_rate(mid, 16000, sr)
    m = min(ref.size, whole.size)
    assert abs(ref.size - whole.size) <= max(sr // 16000, 16000 // sr) * 160 + 8
    assert np.abs(whole[:m] - ref[:m]).max() < 2e-4
    # the lock-step group does the same on [B, n] arrays
    g = StreamGroup(model="dpdfnet2", streams=2)
    xs = np.stack([x, x[::-1].copy()])
    parts = [g.process(xs[:, k:k + 480], sr) for k in range(0, n, 480)]
    parts.append(g.flush())
    gg = np.concatenate(parts, 1)
    assert gg.shape[1] == whole.size and np.abs(gg[0] - whole).max() < 1e-6
    g.close()


def test_device_raised_errors_surface_through_the_abi(random_weights):
    """A non-finite activation in a tensor-core operand converter must fail the hop, not poison the stream silently."""
    from dpdfnet_b200.engine import Engine
    eng = Engine("dpdfnet2", None, max_streams=4)
    for k in ("intra_tc", "post_tc", "sep_tc", "gru_tc", "dft_tc"):
        eng.set_option(k, 1)
    x = np.zeros((4, 160 * 3), np.float32)
    eng.run_pcm_host(x)                                   # digital silence is fine
    x[2, 200] = np.nan
    with pytest.raises(RuntimeError, match="FP16 operand range|non-finite"):
        eng.run_pcm_host(x)
    eng.reset()
    eng.poll_error()                                      # cleared once reported
    eng.run_pcm_host(np.zeros((4, 160), np.float32))
    # a corrupt blob header must be rejected, not read out of bounds (api.cu:parse_blob)
    import struct
    from dpdfnet_b200.spec import get_spec
    bad = b"DPDFW001" + struct.pack("<q", (1 << 62) // 72 + 5) + b"\0" * 256
    with pytest.raises(ValueError):
        Engine(get_spec("dpdfnet2"), bad, max_streams=1)
    eng.close()
