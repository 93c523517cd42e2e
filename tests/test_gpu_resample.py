"""Device resampler against the published polyphase algorithm (scipy.signal.resample_poly), one-shot and streamed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sr_in,sr_out", [(48000, 16000), (16000, 48000), (44100, 16000), (16000, 44100), (8000, 48000)])
def test_batch_resampler_matches_resample_poly(sr_in, sr_out):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    from scipy.signal import resample_poly
    from dpdfnet_b200.resample import BatchResampler
    rng = np.random.default_rng(sr_in + sr_out)
    B, n = 5, 7013
    x = (rng.standard_normal((B, n)) * 0.3).astype(np.float32)
    rs = BatchResampler(sr_in, sr_out, max_streams=B)
    ref = np.stack([resample_poly(x[b].astype(np.float64), rs.up, rs.down) for b in range(B)])
    xd = torch.from_numpy(x).cuda()
    one = rs.resample(xd).cpu().numpy()
    assert one.shape == ref.shape
    assert np.abs(one - ref).max() < 2e-5                               # FP32 taps and accumulation vs float64
    from oracle.resample_np import resample_direct
    from dpdfnet_b200.resample import design_taps
    assert np.abs(one[0] - resample_direct(x[0], rs.up, rs.down, design_taps(rs.up, rs.down))).max() < 2e-5
    # streamed in ragged chunks (including chunks shorter than the filter support) == one shot, bit for bit
    rs.reset()
    parts, pos = [], 0
    for c in [1, 37, 640, 3, 2048, 1111, 1, 999, 5000]:
        c = min(c, n - pos)
        if c <= 0:
            break
        parts.append(rs.process(xd[:, pos:pos + c]))
        pos += c
    assert pos == n
    parts.append(rs.process(xd[:, :0], flush=True))
    streamed = torch.cat(parts, 1).cpu().numpy()
    assert np.array_equal(streamed, one)
    with pytest.raises(ValueError):
        rs.process(torch.zeros(B + 1, 4, device="cuda"))
