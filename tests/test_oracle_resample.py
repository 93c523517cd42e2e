"""The resampler restatement against scipy's published polyphase algorithm (CPU)."""
import numpy as np
import pytest


@pytest.mark.parametrize("sr_in,sr_out", [(48000, 16000), (16000, 48000), (44100, 16000), (16000, 44100)])
def test_direct_sum_equals_resample_poly(sr_in, sr_out):
    from fractions import Fraction
    from scipy.signal import resample_poly
    from dpdfnet_b200.resample import design_taps
    from oracle.resample_np import resample_direct, streamed_count
    fr = Fraction(sr_out, sr_in)
    up, down = fr.numerator, fr.denominator
    taps = design_taps(up, down)
    assert taps.size == 2 * 10 * max(up, down) + 1 and taps.dtype == np.float32
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1500)
    ref = resample_poly(x, up, down)
    got = resample_direct(x, up, down, taps)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-6                      # float32 taps vs scipy's float64 design
    # streaming bookkeeping: counts are monotone, never run ahead of the filter support, and flush completes the clip
    half = taps.size // 2
    prev = 0
    for n in range(0, 1501, 37):
        c = streamed_count(n, up, down, half, False)
        assert prev <= c <= -(-n * up // down)
        if c:
            assert (c - 1) * down + half <= (n - 1) * up
        prev = c
    assert streamed_count(1500, up, down, half, True) == ref.size
