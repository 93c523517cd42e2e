"""Static description of the DPDFNet models served by the engine.

Everything here is derived from the reference constructors; nothing is learned.
Reference: ``onnx_model/dpdfnet.py:522-713`` (16 kHz), ``onnx_model/dpdfnet_48khz_hr.py:586-790``
(48 kHz HR), model registry ``package/src/dpdfnet/models.py:26-69``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

CONV_CH = 64          # conv_ch / DPRNN hidden (onnx_model/dpdfnet.py:534)
GRU_DIM = 256         # enc/erb_dec/df_dec GRU width (:538-540)
EMB_DIM = 512         # conv_ch * nb_erb // 4 (16k) and emb_dim (48k)
NB_DF = 96            # int(4800 / (sr//2) * freq_bins): 96 for both rates (:623)
NB_ERB = 32
DF_ORDER = 5
DF_LOOKAHEAD = 2
CONV_LOOKAHEAD = 2
ALPHA_NORM = 0.98
LATENCY_FRAMES = DF_LOOKAHEAD + CONV_LOOKAHEAD   # 4 frames, see audio.py:8


def vorbis_window(win: int) -> np.ndarray:
    """w[i] = sin(pi/2 * sin^2(pi (i+0.5)/win)); model/utils.py:153-161, audio.py:84-88."""
    i = np.arange(win, dtype=np.float64)
    s = np.sin(0.5 * np.pi * (i + 0.5) / (win / 2))
    return np.sin(0.5 * np.pi * s * s)


def erb_band_widths(nfft: int, sr: int, nb_bands: int, min_nb_freqs: int) -> List[int]:
    """Widths of the rectangular, non-overlapping ERB bands.

    Restates ``model/utils.py:265-324`` (``erb_filter_banks``) including its quirk that band
    edges are only generated for 33 points (so only 32 bands are valid).
    """
    assert nb_bands == 32, "reference only supports 32 ERB bands (model/utils.py:304)"

    def freq2erb(f):
        return 9.265 * np.log1p(f / (24.7 * 9.265))

    def erb2freq(e):
        return 24.7 * 9.265 * (np.exp(e / 9.265) - 1)

    nyq = sr / 2
    fw = sr / nfft
    lo, hi = freq2erb(0.0), freq2erb(nyq)
    step = (hi - lo) / nb_bands
    bins = np.zeros(nb_bands + 1, dtype=np.int64)
    for i in range(33):
        bins[i] = int(round(erb2freq(lo + i * step) / fw))
    bins[-1] = nfft // 2 + 1
    nfreq = nfft // 2 + 1
    member = np.zeros((nb_bands, nfreq), dtype=np.int64)
    over = 0
    for j in range(nb_bands):
        a, b = bins[j] + over, bins[j + 1]
        if (b - a) < min_nb_freqs:
            over = min_nb_freqs - (b - a)
            b = min(b + over, nfreq)
        else:
            over = 0
        member[j, a:b] = 1
    # every bin must belong to exactly one band and bands must be contiguous
    assert (member.sum(0) == 1).all(), "ERB bands are expected to partition the spectrum"
    widths = member.sum(1).tolist()
    start = 0
    for j, w in enumerate(widths):
        assert member[j, start:start + w].all()
        start += w
    return widths


@dataclass(frozen=True)
class ModelSpec:
    name: str
    sample_rate: int
    n_blocks: int                 # dprnn_num_blocks
    hr48: bool                    # DPDFNet48HR variant
    win: int = field(init=False)
    hop: int = field(init=False)
    freq_bins: int = field(init=False)
    fe_feat: int = field(init=False)      # width of the "erb" feature vector (32 | 481)
    fe: Tuple[int, int, int, int] = field(init=False)   # e0..e3 frequency widths
    erb_strides: Tuple[int, int, int] = field(init=False)
    dec_up: Tuple[int, int, int] = field(init=False)    # convt3, convt2, convt1 up-factors (1 = plain conv)

    def __post_init__(self):
        win = int(0.02 * self.sample_rate)
        object.__setattr__(self, "win", win)
        object.__setattr__(self, "hop", win // 2)
        object.__setattr__(self, "freq_bins", win // 2 + 1)
        if self.hr48:
            fe0 = win // 2                       # 480: feat[..., :-1], dpdfnet_48khz_hr.py:263
            strides = (3, 2, 2)
            fe = (fe0, fe0 // 3, fe0 // 6, fe0 // 12)
            object.__setattr__(self, "fe_feat", win // 2 + 1)
            object.__setattr__(self, "dec_up", (2, 2, 3))
        else:
            strides = (2, 2, 1)
            fe = (NB_ERB, NB_ERB // 2, NB_ERB // 4, NB_ERB // 4)
            object.__setattr__(self, "fe_feat", NB_ERB)
            object.__setattr__(self, "dec_up", (1, 2, 2))
        object.__setattr__(self, "fe", fe)
        object.__setattr__(self, "erb_strides", strides)

    # ----- derived sizes -------------------------------------------------
    @property
    def nb_df(self) -> int:
        return NB_DF

    @property
    def fd(self) -> Tuple[int, int]:
        return (NB_DF, NB_DF // 2)

    @property
    def wnorm(self) -> float:
        """model/utils.py:164-167."""
        return 1.0 / (self.win ** 2 / (2 * self.hop))

    @property
    def erb_widths(self) -> List[int]:
        return erb_band_widths(self.win, self.sample_rate, NB_ERB, 2 if self.hr48 else 1)

    # ----- reference flat state layout (onnx_model/dpdfnet.py:737-746) ----
    def state_segments(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """Ordered (name, shape) of the reference's flat state vector."""
        F = self.freq_bins
        N = self.n_blocks
        segs: List[Tuple[str, Tuple[int, ...]]] = [
            ("erb_norm.mu", (self.fe_feat,)),
            ("spec_norm.s", (NB_DF,)),
            ("enc.erb_conv0.ring", (3, self.fe_feat)),
        ]
        segs += [(f"enc.dprnn_erb.{i}.h", (self.fe[3], CONV_CH)) for i in range(N)]
        segs += [("enc.df_conv0.ring", (3, 2, NB_DF))]
        segs += [(f"enc.dprnn_df.{i}.h", (NB_DF // 2, CONV_CH)) for i in range(N)]
        segs += [
            ("enc.emb_gru.h", (GRU_DIM,)),
            ("erb_dec.emb_gru.h", (2, GRU_DIM)),
            ("df_dec.df_gru.h", (2, GRU_DIM)),
            ("df_dec.c0.ring", (DF_ORDER, CONV_CH, NB_DF)),
            ("mask.ring", (3, F, 2)),
            ("df_op.coef.ring", (3, DF_ORDER, NB_DF, 2)),
            ("df_op.spec.ring", (DF_ORDER, F, 2)),
        ]
        return segs

    @property
    def state_size(self) -> int:
        return int(sum(int(np.prod(s)) for _, s in self.state_segments()))

    # ----- roofline accounting ----------------------
    @property
    def state_write_floats(self) -> int:
        """Floats of state that change per frame: norms + one slot per ring + all GRU states."""
        F = self.freq_bins
        N = self.n_blocks
        return (self.fe_feat + NB_DF) + (self.fe_feat + 2 * NB_DF) \
            + CONV_CH * (self.fe[3] + NB_DF // 2) * N + GRU_DIM * 5 \
            + CONV_CH * NB_DF + 2 * F + 2 * DF_ORDER * NB_DF + 2 * F

    @property
    def algorithmic_bytes_per_frame(self) -> int:
        """4*(2*hop) + 4*(S + S_write): BASELINE.md section 2."""
        return 4 * (2 * self.hop) + 4 * (self.state_size + self.state_write_floats)

    @property
    def macs_per_frame(self) -> int:
        """Multiply-accumulates of the network per stream-frame (dense DFTs excluded)."""
        C = CONV_CH
        fe0, fe1, fe2, fe3 = self.fe
        macs = fe0 * C * 9                                   # erb_conv0
        macs += (fe1 + fe2 + fe3) * (3 * C + C * C)          # erb_conv1..3
        macs += NB_DF * (C * 9 + C * C)                      # df_conv0
        macs += (NB_DF // 2) * (3 * C + C * C)               # df_conv1
        per_pos = 2 * (2 * C * 3 * C) + 2 * C * C + 2 * C * 3 * C + C * C
        macs += self.n_blocks * per_pos * (fe3 + NB_DF // 2)  # DPRNN
        macs += C * (NB_DF // 2) * 16                        # df_fc_emb 3072->512, 32 groups
        if self.hr48:
            macs += 2 * (C * fe3) * 16                       # erb_fc_emb enc + dec
        macs += 1024 * 16 + 256 * 32                         # enc GL in/out
        macs += 5 * 2 * GRU_DIM * 3 * GRU_DIM                # five GRUCell(256)
        macs += 512 * 16 + 256 * 32                          # erb_dec GL in/out
        macs += 512 * 32 + 512 * 16 + 256 * 60               # df_gru in (8g), df_skip, df_out
        up3, up2, up1 = self.dec_up
        macs += (fe3 * up3) * (3 * C + C * C)                # convt3
        macs += fe1 * (3 * C + C * C) + fe0 * (3 * C + C * C)  # convt2, convt1
        macs += (fe0 + fe1 + fe2 + fe3) * C                  # pathway convs
        macs += fe0 * 3 * C                                  # conv0_out
        macs += NB_DF * (10 * 5 * 32 + 100)                  # df_convp
        macs += NB_DF * DF_ORDER * 4 + self.freq_bins * 2    # deep filter + mask
        return int(macs)


MODEL_SPECS: Dict[str, ModelSpec] = {
    "baseline": ModelSpec("baseline", 16000, 0, False),
    "dpdfnet2": ModelSpec("dpdfnet2", 16000, 2, False),
    "dpdfnet4": ModelSpec("dpdfnet4", 16000, 4, False),
    "dpdfnet8": ModelSpec("dpdfnet8", 16000, 8, False),
    "dpdfnet2_48khz_hr": ModelSpec("dpdfnet2_48khz_hr", 48000, 2, True),
    "dpdfnet8_48khz_hr": ModelSpec("dpdfnet8_48khz_hr", 48000, 8, True),
}


def get_spec(name: str) -> ModelSpec:
    try:
        return MODEL_SPECS[name]
    except KeyError:
        raise ValueError(f"Unknown model {name!r}; available: {sorted(MODEL_SPECS)}") from None
