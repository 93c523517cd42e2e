"""Checkpoint ingestion: reference ``state_dict`` -> packed device blob.

The reference ships plain ``state_dict`` ``.pth`` files in the *offline* naming
(``model/dpdfnet.py:642-643``); the streaming model renames a few keys
(``onnx_model/dpdfnet.py:876-888``) but the arithmetic is the same.  This module

* enumerates the parameter shapes of a checkpoint (``ref_param_shapes``),
* builds a seeded random checkpoint with **randomised BatchNorm running statistics**
  (no shipped weights are available offline; fresh BN would hide folding bugs),
* packs a checkpoint into the engine's blob: eval-BatchNorm folded into the preceding
  bias-free convolution (eps 1e-5, ``torch.nn.BatchNorm2d`` default), GRU biases pre-summed
  for the r/z gates, grouped linears stacked ``[G, O/G, I/G]``, and the windowed DFT bases.

Blob format (little endian): ``b"DPDFW001"``, ``int64 n_entries``, then ``n_entries`` records of
``char name[56]; int64 offset_floats; int64 numel`` followed by the float32 payload.  Offsets
are relative to the payload start and 32-float (128 B) aligned.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, Mapping, Tuple

import numpy as np

from .spec import CONV_CH, DF_ORDER, GRU_DIM, NB_DF, ModelSpec, vorbis_window

MAGIC = b"DPDFW001"
NAME_LEN = 56
ALIGN = 32  # floats
LOG2E = 1.4426950408889634
BN_EPS = 1e-5


# --------------------------------------------------------------------------
# reference checkpoint shape table
# --------------------------------------------------------------------------

def _gl(shapes, prefix: str, groups: int, k: int, n: int):
    for g in range(groups):
        shapes[f"{prefix}.layers.{g}.weight"] = (n, k)
        shapes[f"{prefix}.layers.{g}.bias"] = (n,)


def _bn(shapes, prefix: str, ch: int):
    shapes[f"{prefix}.weight"] = (ch,)
    shapes[f"{prefix}.bias"] = (ch,)
    shapes[f"{prefix}.running_mean"] = (ch,)
    shapes[f"{prefix}.running_var"] = (ch,)


def _gru(shapes, prefix: str, inp: int, hid: int, layers: int = 1, reverse: bool = False):
    for l in range(layers):
        for suf in ([""] + (["_reverse"] if reverse else [])):
            shapes[f"{prefix}.weight_ih_l{l}{suf}"] = (3 * hid, inp if l == 0 else hid)
            shapes[f"{prefix}.weight_hh_l{l}{suf}"] = (3 * hid, hid)
            shapes[f"{prefix}.bias_ih_l{l}{suf}"] = (3 * hid,)
            shapes[f"{prefix}.bias_hh_l{l}{suf}"] = (3 * hid,)


def ref_param_shapes(spec: ModelSpec) -> "OrderedDict[str, Tuple[int, ...]]":
    """Learned tensors of a reference checkpoint (offline naming), buffers excluded."""
    C = CONV_CH
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["enc.erb_conv0.1.weight"] = (C, 1, 3, 3)
    _bn(s, "enc.erb_conv0.2", C)
    for i in (1, 2, 3):
        s[f"enc.erb_conv{i}.0.weight"] = (C, 1, 1, 3)
        s[f"enc.erb_conv{i}.1.weight"] = (C, C, 1, 1)
        _bn(s, f"enc.erb_conv{i}.2", C)
    s["enc.df_conv0.1.convs.0.weight"] = (C // 2, 1, 3, 3)
    s["enc.df_conv0.1.convs.1.weight"] = (C // 2, 1, 3, 3)
    s["enc.df_conv0.2.weight"] = (C, C, 1, 1)
    _bn(s, "enc.df_conv0.3", C)
    s["enc.df_conv1.0.weight"] = (C, 1, 1, 3)
    s["enc.df_conv1.1.weight"] = (C, C, 1, 1)
    _bn(s, "enc.df_conv1.2", C)
    for br in ("erb", "df"):
        for i in range(spec.n_blocks):
            p = f"enc.dprnn_{br}.blocks.{i}"
            _gru(s, f"{p}.intra_gru", C, C, reverse=True)
            s[f"{p}.fc_intra.weight"] = (C, 2 * C)
            s[f"{p}.fc_intra.bias"] = (C,)
            s[f"{p}.ln_intra.weight"] = (C,)
            s[f"{p}.ln_intra.bias"] = (C,)
            _gru(s, f"{p}.inter_gru", C, C)
            s[f"{p}.fc_inter.weight"] = (C, C)
            s[f"{p}.fc_inter.bias"] = (C,)
            s[f"{p}.ln_inter.weight"] = (C,)
            s[f"{p}.ln_inter.bias"] = (C,)
    if spec.hr48:
        _gl(s, "enc.erb_fc_emb.0", 32, C * spec.fe[3] // 32, 16)
    _gl(s, "enc.df_fc_emb.0", 32, C * (NB_DF // 2) // 32, 16)
    _gl(s, "enc.emb_gru.linear_in.0", 16, 64, 16)
    _gru(s, "enc.emb_gru.gru", GRU_DIM, GRU_DIM)
    _gl(s, "enc.emb_gru.linear_out.0", 16, 16, 32)
    s["enc.lsnr_fc.0.weight"] = (1, 512)
    s["enc.lsnr_fc.0.bias"] = (1,)
    _gl(s, "erb_dec.emb_gru.linear_in.0", 16, 32, 16)
    _gru(s, "erb_dec.emb_gru.gru", GRU_DIM, GRU_DIM, layers=2)
    _gl(s, "erb_dec.emb_gru.linear_out.0", 16, 16, 32)
    if spec.hr48:
        _gl(s, "erb_dec.erb_fc_emb.0", 32, 16, C * spec.fe[3] // 32)
    up3, up2, up1 = spec.dec_up
    for i, up in ((3, up3), (2, up2), (1, up1)):
        s[f"erb_dec.conv{i}p.0.weight"] = (C, 1, 1, 1)
        _bn(s, f"erb_dec.conv{i}p.1", C)
        if up == 1:
            s[f"erb_dec.convt{i}.0.weight"] = (C, 1, 1, 3)
        else:
            for j in range(up):
                s[f"erb_dec.convt{i}.0.convs.{j}.weight"] = (C, 1, 1, 3)
        s[f"erb_dec.convt{i}.1.weight"] = (C, C, 1, 1)
        _bn(s, f"erb_dec.convt{i}.2", C)
    s["erb_dec.conv0p.0.weight"] = (C, 1, 1, 1)
    _bn(s, "erb_dec.conv0p.1", C)
    s["erb_dec.conv0_out.0.weight"] = (1, C, 1, 3)
    _bn(s, "erb_dec.conv0_out.1", 1)
    s["df_dec.df_convp.1.convs.0.weight"] = (DF_ORDER, C // 2, DF_ORDER, 1)
    s["df_dec.df_convp.1.convs.1.weight"] = (DF_ORDER, C // 2, DF_ORDER, 1)
    s["df_dec.df_convp.2.weight"] = (2 * DF_ORDER, 2 * DF_ORDER, 1, 1)
    _bn(s, "df_dec.df_convp.3", 2 * DF_ORDER)
    _gl(s, "df_dec.df_gru.linear_in.0", 8, 64, 32)       # default linear_groups=8 quirk
    _gru(s, "df_dec.df_gru.gru", GRU_DIM, GRU_DIM, layers=2)
    _gl(s, "df_dec.df_skip", 16, 32, 16)
    _gl(s, "df_dec.df_out.0", 16, 16, NB_DF * 2 * DF_ORDER // 16)
    return s


def random_checkpoint(spec: ModelSpec, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Deterministic (numpy PCG64) stand-in for a shipped checkpoint, reference naming."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in ref_param_shapes(spec).items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_var":
            v = rng.uniform(0.5, 1.5, shape)
        elif leaf == "running_mean":
            v = rng.uniform(-0.2, 0.2, shape)
        elif leaf == "weight" and len(shape) == 1:
            v = rng.uniform(0.5, 1.5, shape)          # LayerNorm / BatchNorm gamma
        elif len(shape) == 1:
            v = rng.uniform(-0.2, 0.2, shape)          # all biases (incl. BN/LN beta, GRU biases)
        else:
            fan_in = int(np.prod(shape[1:]))
            if "gru" in name and "linear" not in name:
                bound = np.sqrt(6.0 / (shape[0] // 3 + shape[1]))   # ~xavier per gate
            else:
                bound = np.sqrt(3.0 / fan_in)
            v = rng.uniform(-bound, bound, shape)
        sd[name] = np.ascontiguousarray(v, dtype=np.float32)
    return sd


# --------------------------------------------------------------------------
# packing
# --------------------------------------------------------------------------

def _bn_affine(sd: Mapping[str, np.ndarray], prefix: str) -> Tuple[np.ndarray, np.ndarray]:
    g = sd[f"{prefix}.weight"].astype(np.float64)
    b = sd[f"{prefix}.bias"].astype(np.float64)
    m = sd[f"{prefix}.running_mean"].astype(np.float64)
    v = sd[f"{prefix}.running_var"].astype(np.float64)
    sc = g / np.sqrt(v + BN_EPS)
    return sc, b - m * sc


def _gl_pack(sd, prefix: str, groups: int) -> Tuple[np.ndarray, np.ndarray]:
    w = np.stack([sd[f"{prefix}.layers.{g}.weight"] for g in range(groups)], 0)   # [G, O/G, I/G]
    b = np.concatenate([sd[f"{prefix}.layers.{g}.bias"] for g in range(groups)], 0)
    return w, b


def _gru_bias(b_ih: np.ndarray, b_hh: np.ndarray) -> np.ndarray:
    """[4, H]: (b_ir+b_hr, b_iz+b_hz, b_in, b_hn); gate order r,z,n as torch.nn.GRU."""
    H = b_ih.shape[0] // 3
    bi, bh = b_ih.reshape(3, H).astype(np.float64), b_hh.reshape(3, H).astype(np.float64)
    return np.stack([bi[0] + bh[0], bi[1] + bh[1], bi[2], bh[2]], 0)


def dft_bases(spec: ModelSpec) -> Dict[str, np.ndarray]:
    """Windowed real-DFT bases with ``wnorm`` folded in.

    Analysis (stream.py:119-126 / model/dpdfnet.py:608-613): ``X_k = wnorm * sum_n x_n w_n e^{-2 pi i nk/N}``.
    Synthesis (stream.py:138-144 / model/dpdfnet.py:615-625): ``y_n = w_n / wnorm * irfft(Y)_n``.
    """
    N, F = spec.win, spec.freq_bins
    w = vorbis_window(N)
    n = np.arange(N, dtype=np.float64)[:, None]
    k = np.arange(F, dtype=np.float64)[None, :]
    ang = 2.0 * np.pi * ((n * k) % N) / N
    fwd_c = (w[:, None] * np.cos(ang)) * spec.wnorm            # [N, F]
    fwd_s = (-w[:, None] * np.sin(ang)) * spec.wnorm
    ck = np.full(F, 2.0)
    ck[0] = 1.0
    ck[-1] = 1.0
    inv_c = (ck[:, None] * np.cos(ang.T)) * (w[None, :] / (N * spec.wnorm))     # [F, N]
    inv_s = (-ck[:, None] * np.sin(ang.T)) * (w[None, :] / (N * spec.wnorm))
    inv_s[0, :] = 0.0
    inv_s[-1, :] = 0.0
    # interleaved (cos, sin) pairs: one 64-bit load feeds one packed FFMA2 on the device
    return {"const.dft_fwd": np.stack([fwd_c, fwd_s], -1),      # [N, F, 2]
            "const.dft_inv": np.stack([inv_c, inv_s], -1)}      # [F, N, 2]


DFT_TC_NC = 128      # output columns (re / im interleaved) per CTA of the analysis GEMM
IDFT_TC_W = 80       # output samples of each frame half per CTA of the synthesis GEMM (NC = 2 * 80)


def dft_tc_images(spec: ModelSpec, bases: Mapping[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """FP16 hi | lo tcgen05 operand images of the two DFT bases for k_dft_tc (csrc/k_dft_tc.cu): the framed DFT and the
    inverse DFT + overlap-add as [128 streams x K] x [K x N] GEMMs.

    Analysis: N = 2F columns (re_0, im_0, re_1, ...), padded with zero rows to a multiple of DFT_TC_NC; K = win in
    64-wide stages.  Layout [chunk][stage][NC rows x 64 k] (hi image, lo image).
    Synthesis: K = 2F (re_0, im_0, ...) zero-padded to a multiple of 64; a chunk owns IDFT_TC_W samples n of the first
    frame half AND the samples n + hop of the second, so one CTA reads the old overlap-add tail, emits the hop and writes
    the new tail of the same addresses.  Layout [chunk][stage][2 W rows x 64 k].

    Scaling.  The hi / lo split is only FP32-accurate while ``lo`` stays a NORMAL half (|x| above ~1e-4 x 2^11): the
    analysis basis peaks at wnorm = 1/320 (1/960) and -120 dBFS audio at 1e-6, both deep in the subnormals.  So the
    basis is stored times a power of two that brings its peak to [0.5, 1), the kernel multiplies PCM (and Y) by 2^13
    before splitting (full scale stays below the FP16 maximum up to |x| = 8) and the accumulators are scaled back
    exactly: ``const.dft_tc_scale`` = (in, out) factors of the analysis, then of the synthesis."""
    N, F, hop = spec.win, spec.freq_bins, spec.hop
    fwd = np.asarray(bases["const.dft_fwd"], dtype=np.float64).reshape(N, 2 * F)      # [n][2 bin + ri]
    inv = np.asarray(bases["const.dft_inv"], dtype=np.float64)                         # [bin][n][ri]
    assert N % 64 == 0 and hop % IDFT_TC_W == 0
    ncol = -(-2 * F // DFT_TC_NC) * DFT_TC_NC
    s_in = 2.0 ** 13
    s_bf = 2.0 ** np.floor(-np.log2(np.abs(fwd).max()))        # peak of the scaled basis in [0.5, 1)
    s_bi = 2.0 ** np.floor(-np.log2(np.abs(inv).max()))
    fwd_p = np.zeros((ncol, N), dtype=np.float32)
    fwd_p[:2 * F] = fwd.T * s_bf
    out_f = [umma_operand16(np.ascontiguousarray(fwd_p[c:c + DFT_TC_NC, k:k + 64]))
             for c in range(0, ncol, DFT_TC_NC) for k in range(0, N, 64)]
    kpad = -(-2 * F // 64) * 64
    inv_p = np.zeros((N, kpad), dtype=np.float32)                                      # [n][2 bin + ri]
    inv_p[:, :2 * F] = inv.transpose(1, 0, 2).reshape(N, 2 * F) * s_bi
    out_i = []
    for c in range(hop // IDFT_TC_W):
        rows = np.concatenate([np.arange(c * IDFT_TC_W, (c + 1) * IDFT_TC_W), hop + np.arange(c * IDFT_TC_W, (c + 1) * IDFT_TC_W)])
        for k in range(0, kpad, 64):
            out_i.append(umma_operand16(np.ascontiguousarray(inv_p[rows, k:k + 64])))
    scale = np.array([s_in, 1.0 / (s_in * s_bf), s_in, 1.0 / (s_in * s_bi)], dtype=np.float32)
    return {"const.dft_fwd_tc": np.concatenate(out_f), "const.dft_inv_tc": np.concatenate(out_i), "const.dft_tc_scale": scale}


def norm_init(spec: ModelSpec) -> Tuple[np.ndarray, np.ndarray]:
    """(mu0, s0): ErbNorm/SpecNorm linspace inits (layers.py:460-463, 519-522) or the 48 kHz tables."""
    if spec.hr48:
        from .data import norm_init_48k
        return norm_init_48k.MU0.astype(np.float32), norm_init_48k.S0.astype(np.float32)
    # float32 arithmetic exactly as torch: init0 + arange * step
    step_mu = np.float32((-90.0 - -60.0) / (spec.fe_feat - 1))
    mu0 = np.float32(-60.0) + np.arange(spec.fe_feat, dtype=np.float32) * step_mu
    step_s = np.float32((0.0001 - 0.001) / (NB_DF - 1))
    s0 = np.float32(0.001) + np.arange(NB_DF, dtype=np.float32) * step_s
    return mu0.astype(np.float32), s0.astype(np.float32)


def pack_tensors(spec: ModelSpec, sd: Mapping[str, np.ndarray]) -> "OrderedDict[str, np.ndarray]":
    """Reference checkpoint -> named float32 tensors in engine layout."""
    C = CONV_CH
    sd = {k: np.asarray(v) for k, v in sd.items()}
    # enc.lsnr_fc is computed by the reference but never leaves the ONNX graph: the engine does not use it and
    # an .onnx export does not contain it, so it is optional here
    want = {k: shp for k, shp in ref_param_shapes(spec).items() if ".lsnr_fc." not in k or k in sd}
    missing = [k for k in want if k not in sd]
    if missing:
        raise KeyError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
    for k, shp in want.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError(f"{k}: expected shape {shp}, got {tuple(sd[k].shape)}")
    t: "OrderedDict[str, np.ndarray]" = OrderedDict()
    bases = dft_bases(spec)
    for k, v in bases.items():
        t[k] = v
    mu0, s0 = norm_init(spec)
    t["const.mu0"], t["const.s0"] = mu0, s0
    for k, v in dft_tc_images(spec, bases).items():
        t[k] = v

    def sep(prefix_out: str, dw_keys, pw_key: str, bn_prefix: str):
        # depthwise [S][3][C] (tap-major, channel contiguous); pointwise [C_out][C_in] with BN scale folded
        dws = [sd[k][:, 0, 0, :].T for k in dw_keys]                       # [3, C] each
        sc, sh = _bn_affine(sd, bn_prefix)
        t[f"{prefix_out}.dw"] = np.stack(dws, 0)
        t[f"{prefix_out}.pw"] = sd[pw_key][:, :, 0, 0].astype(np.float64) * sc[:, None]
        t[f"{prefix_out}.tc_pw"] = umma_operand16(t[f"{prefix_out}.pw"].astype(np.float32))     # k_sepconv_tc
        t[f"{prefix_out}.b"] = sh

    # --- encoder -----------------------------------------------------------
    sc, sh = _bn_affine(sd, "enc.erb_conv0.2")
    w = sd["enc.erb_conv0.1.weight"][:, 0].astype(np.float64) * sc[:, None, None]       # [C,3,3]
    t["enc.erb_conv0.w"] = w.reshape(C, 9).T                                            # [9, C]
    t["enc.erb_conv0.b"] = sh
    for i in (1, 2, 3):
        sep(f"enc.erb_conv{i}", [f"enc.erb_conv{i}.0.weight"], f"enc.erb_conv{i}.1.weight", f"enc.erb_conv{i}.2")
    gw = np.concatenate([sd["enc.df_conv0.1.convs.0.weight"][:, 0], sd["enc.df_conv0.1.convs.1.weight"][:, 0]], 0)
    t["enc.df_conv0.w"] = gw.reshape(C, 9).T                                            # [9, C]; ch<32 <- re, else im
    sc, sh = _bn_affine(sd, "enc.df_conv0.3")
    t["enc.df_conv0.pw"] = sd["enc.df_conv0.2.weight"][:, :, 0, 0].astype(np.float64) * sc[:, None]
    t["enc.df_conv0.tc_pw"] = umma_operand16(t["enc.df_conv0.pw"].astype(np.float32))
    t["enc.df_conv0.b"] = sh
    sep("enc.df_conv1", ["enc.df_conv1.0.weight"], "enc.df_conv1.1.weight", "enc.df_conv1.2")

    for br in ("erb", "df"):
        for i in range(spec.n_blocks):
            p = f"enc.dprnn_{br}.blocks.{i}"
            q = f"enc.dprnn_{br}.{i}"
            t[f"{q}.intra.wih"] = np.stack([sd[f"{p}.intra_gru.weight_ih_l0"], sd[f"{p}.intra_gru.weight_ih_l0_reverse"]], 0)
            t[f"{q}.intra.whh"] = np.stack([sd[f"{p}.intra_gru.weight_hh_l0"], sd[f"{p}.intra_gru.weight_hh_l0_reverse"]], 0)
            t[f"{q}.intra.bias"] = np.stack([
                _gru_bias(sd[f"{p}.intra_gru.bias_ih_l0"], sd[f"{p}.intra_gru.bias_hh_l0"]),
                _gru_bias(sd[f"{p}.intra_gru.bias_ih_l0_reverse"], sd[f"{p}.intra_gru.bias_hh_l0_reverse"])], 0)
            t[f"{q}.intra.fc_w"] = sd[f"{p}.fc_intra.weight"]
            t[f"{q}.intra.fc_b"] = sd[f"{p}.fc_intra.bias"]
            t[f"{q}.intra.ln_g"] = sd[f"{p}.ln_intra.weight"]
            t[f"{q}.intra.ln_b"] = sd[f"{p}.ln_intra.bias"]
            t[f"{q}.inter.wih"] = sd[f"{p}.inter_gru.weight_ih_l0"]
            t[f"{q}.inter.whh"] = sd[f"{p}.inter_gru.weight_hh_l0"]
            t[f"{q}.inter.bias"] = _gru_bias(sd[f"{p}.inter_gru.bias_ih_l0"], sd[f"{p}.inter_gru.bias_hh_l0"])
            t[f"{q}.inter.fc_w"] = sd[f"{p}.fc_inter.weight"]
            t[f"{q}.inter.fc_b"] = sd[f"{p}.fc_inter.bias"]
            t[f"{q}.inter.ln_g"] = sd[f"{p}.ln_inter.weight"]
            t[f"{q}.inter.ln_b"] = sd[f"{p}.ln_inter.bias"]
            # tensor-core (tcgen05, FP16 hi/lo split) operand images of the position-parallel matrices of the block:
            # nine [64x64] slabs (hi | lo, 16 KB each): fc_intra K-halves, inter-GRU gates (Wih r,z,n then Whh r,z,n), fc_inter
            wih, whh = sd[f"{p}.inter_gru.weight_ih_l0"], sd[f"{p}.inter_gru.weight_hh_l0"]
            fcw = sd[f"{p}.fc_intra.weight"]
            t[f"{q}.tc.fc_w"] = np.concatenate([umma_operand16(fcw[:, :C]), umma_operand16(fcw[:, C:])])   # two K=64 slabs
            t[f"{q}.tc.gates"] = np.concatenate([umma_operand16(m[g * C:(g + 1) * C]) for m in (wih, whh) for g in range(3)])
            t[f"{q}.tc.fc2_w"] = umma_operand16(sd[f"{p}.fc_inter.weight"])
            # FP16 hi/lo operand images of the intra-frame GRU (k_dprnn_intra_tc): per direction
            # [W_ih hi | W_ih lo | W_hh hi | W_hh lo], each [192][64] halves stored as raw bytes in the f32 blob.
            # The exponent scales of the gate non-linearities are folded into rows and biases: the kernel
            # evaluates sigmoid(a) = 1 / (1 + 2^(-log2(e) a)) and tanh(c) = 1 - 2 / (1 + 2^(2 log2(e) c)).
            gscale = np.repeat(np.array([-LOG2E, -LOG2E, 2.0 * LOG2E]), C).astype(np.float32)[:, None]
            t[f"{q}.tc.intra"] = np.concatenate([
                umma_operand16(sd[f"{p}.intra_gru.{m}_l0{sfx}"] * gscale) for sfx in ("", "_reverse") for m in ("weight_ih", "weight_hh")])
            if True:
                # fragment form of the sweep (k_dprnn_intra_tc.cu:intra_sweep_f): the thread of unit (cg, j) packs its units
                # of K slices 2p and 2p + 1 into ONE operand column 16 p + 4 cg + j, so K element k = 2 c + e of the
                # recurrent product is hidden unit 16 (2 p + e) + 4 cg + j; W_ih is unchanged
                kk = np.arange(C)
                cc, ee = kk >> 1, kk & 1
                perm = 16 * (2 * (cc >> 4) + ee) + 4 * ((cc >> 2) & 3) + (cc & 3)
                assert sorted(perm.tolist()) == list(range(C))
                t[f"{q}.tc.intra_f"] = np.concatenate([
                    umma_operand16((sd[f"{p}.intra_gru.{m}_l0{sfx}"] * gscale)[:, perm if m == "weight_hh" else kk])
                    for sfx in ("", "_reverse") for m in ("weight_ih", "weight_hh")])
            t[f"{q}.tc.intra_bias"] = t[f"{q}.intra.bias"] * np.repeat(
                np.array([-LOG2E, -LOG2E, 2.0 * LOG2E, 2.0 * LOG2E]), C).astype(np.float32).reshape(1, 4, C)

    def gl(out: str, prefix: str, groups: int):
        t[f"{out}.w"], t[f"{out}.b"] = _gl_pack(sd, prefix, groups)

    def gru(out: str, prefix: str, layers: int):
        for l in range(layers):
            t[f"{out}.{l}.wih"] = sd[f"{prefix}.weight_ih_l{l}"]
            t[f"{out}.{l}.whh"] = sd[f"{prefix}.weight_hh_l{l}"]
            t[f"{out}.{l}.bias"] = _gru_bias(sd[f"{prefix}.bias_ih_l{l}"], sd[f"{prefix}.bias_hh_l{l}"])
            t[f"{out}.{l}.tc_w"] = gru_tc_images(t[f"{out}.{l}.wih"], t[f"{out}.{l}.whh"])

    if spec.hr48:
        gl("enc.erb_fc_emb", "enc.erb_fc_emb.0", 32)
    gl("enc.df_fc_emb", "enc.df_fc_emb.0", 32)
    gl("enc.emb_gru.lin_in", "enc.emb_gru.linear_in.0", 16)
    gru("enc.emb_gru.gru", "enc.emb_gru.gru", 1)
    gl("enc.emb_gru.lin_out", "enc.emb_gru.linear_out.0", 16)

    # --- ERB decoder -------------------------------------------------------
    gl("erb_dec.emb_gru.lin_in", "erb_dec.emb_gru.linear_in.0", 16)
    gru("erb_dec.emb_gru.gru", "erb_dec.emb_gru.gru", 2)
    gl("erb_dec.emb_gru.lin_out", "erb_dec.emb_gru.linear_out.0", 16)
    if spec.hr48:
        gl("erb_dec.erb_fc_emb", "erb_dec.erb_fc_emb.0", 32)
    for i, up in zip((3, 2, 1), spec.dec_up):
        sc, sh = _bn_affine(sd, f"erb_dec.conv{i}p.1")
        t[f"erb_dec.conv{i}p.a"] = sd[f"erb_dec.conv{i}p.0.weight"].reshape(C).astype(np.float64) * sc
        t[f"erb_dec.conv{i}p.b"] = sh
        keys = [f"erb_dec.convt{i}.0.weight"] if up == 1 else [f"erb_dec.convt{i}.0.convs.{j}.weight" for j in range(up)]
        sep(f"erb_dec.convt{i}", keys, f"erb_dec.convt{i}.1.weight", f"erb_dec.convt{i}.2")
    sc, sh = _bn_affine(sd, "erb_dec.conv0p.1")
    t["erb_dec.conv0p.a"] = sd["erb_dec.conv0p.0.weight"].reshape(C).astype(np.float64) * sc
    t["erb_dec.conv0p.b"] = sh
    sc, sh = _bn_affine(sd, "erb_dec.conv0_out.1")
    t["erb_dec.conv0_out.w"] = sd["erb_dec.conv0_out.0.weight"][0, :, 0, :].T.astype(np.float64) * sc[0]   # [3, C]
    t["erb_dec.conv0_out.b"] = sh

    # --- DF decoder --------------------------------------------------------
    gw = np.concatenate([sd["df_dec.df_convp.1.convs.0.weight"][..., 0],
                         sd["df_dec.df_convp.1.convs.1.weight"][..., 0]], 0)            # [10, 32, 5(kt)]
    t["df_dec.df_convp.w"] = gw.transpose(0, 2, 1)                                       # [10, 5(kt), 32]
    sc, sh = _bn_affine(sd, "df_dec.df_convp.3")
    t["df_dec.df_convp.pw"] = sd["df_dec.df_convp.2.weight"][:, :, 0, 0].astype(np.float64) * sc[:, None]
    t["df_dec.df_convp.b"] = sh
    gl("df_dec.df_gru.lin_in", "df_dec.df_gru.linear_in.0", 8)
    gru("df_dec.df_gru.gru", "df_dec.df_gru.gru", 2)
    gl("df_dec.df_skip", "df_dec.df_skip", 16)
    gl("df_dec.df_out", "df_dec.df_out.0", 16)
    return OrderedDict((k, np.ascontiguousarray(v, dtype=np.float32)) for k, v in t.items())


def tf32_split(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """w = hi + lo (+ O(2^-24 |w|)) with hi, lo exactly representable in TF32 (round-to-nearest, ties away:
    ``cvt.rna.tf32.f32``).  Three tensor-core passes a_hi*b_hi + a_lo*b_hi + a_hi*b_lo then reproduce the
    FP32 product to ~2^-22 relative ("3xTF32")."""
    def rna(x):
        u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
        return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    hi = rna(w)
    return hi, rna(w - hi)


def umma_kmajor(mat: np.ndarray) -> np.ndarray:
    """[rows, K] -> tcgen05 K-major SWIZZLE_NONE operand image: 8x4-float core matrices (128 B), core
    matrices adjacent in K contiguous (LBO = 128 B), 8-row groups (K/4)*128 B apart (SBO)."""
    n, k = mat.shape
    assert n % 8 == 0 and k % 4 == 0
    return np.ascontiguousarray(mat.reshape(n // 8, 8, k // 4, 4).transpose(0, 2, 1, 3)).reshape(-1)


def umma_operand(mat: np.ndarray) -> np.ndarray:
    """hi image followed by lo image of a weight matrix [N, K] (B operand of D = A * W^T)."""
    hi, lo = tf32_split(mat)
    return np.concatenate([umma_kmajor(hi), umma_kmajor(lo)])


def fp16_split(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """w = hi + lo with hi, lo in IEEE half precision (round-to-nearest-even, ``cvt.rn.f16.f32``): the same
    11-bit significands as the TF32 split, so hi*hi + lo*hi + hi*lo is again FP32-accurate (~2^-22) as long as
    |w| stays inside the FP16 range (checked)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    if not np.all(np.abs(w) < 6.0e4):
        raise ValueError("weight magnitude outside the FP16 range: the FP16-split tensor-core path cannot represent it")
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def umma_kmajor16(mat: np.ndarray) -> np.ndarray:
    """FP16 [rows, K] -> tcgen05 K-major SWIZZLE_NONE operand image: core matrices of 8 rows x 8 halves (128 B),
    adjacent in K contiguous (LBO = 128 B), 8-row groups (K/8)*128 B apart (SBO)."""
    n, k = mat.shape
    assert n % 8 == 0 and k % 8 == 0 and mat.dtype == np.float16
    return np.ascontiguousarray(mat.reshape(n // 8, 8, k // 8, 8).transpose(0, 2, 1, 3)).reshape(-1)


def umma_operand16(mat: np.ndarray) -> np.ndarray:
    """hi image then lo image of a weight matrix [N, K] in FP16, returned as float32 words (raw bytes)."""
    hi, lo = fp16_split(mat)
    return np.concatenate([umma_kmajor16(hi), umma_kmajor16(lo)]).view(np.float32)


def gru_tc_images(wih: np.ndarray, whh: np.ndarray) -> np.ndarray:
    """Weight slabs of k_gru_tc: for each chunk of 64 hidden units, for W_ih then W_hh, for each 64-wide K chunk, the
    [192][64] matrix of the chunk's r, z and n gate rows as FP16 hi | lo operand images (48 KB per slab)."""
    Hh = wih.shape[1]
    out = []
    for u in range(Hh // 64):
        rows = np.concatenate([np.arange(g * Hh + 64 * u, g * Hh + 64 * u + 64) for g in range(3)])
        for m in (wih, whh):
            for kc in range(Hh // 64):
                out.append(umma_operand16(np.ascontiguousarray(m[rows, 64 * kc:64 * kc + 64], dtype=np.float32)))
    return np.concatenate(out)


def serialize(tensors: Mapping[str, np.ndarray]) -> bytes:
    names = list(tensors)
    offsets, cur = [], 0
    for n in names:
        offsets.append(cur)
        cur += (tensors[n].size + ALIGN - 1) // ALIGN * ALIGN
    head = bytearray(MAGIC + struct.pack("<q", len(names)))
    for n, off in zip(names, offsets):
        nb = n.encode()
        if len(nb) >= NAME_LEN:
            raise ValueError(f"tensor name too long: {n}")
        head += nb.ljust(NAME_LEN, b"\0") + struct.pack("<qq", off, tensors[n].size)
    pad = (-len(head)) % 128
    head += b"\0" * pad
    payload = np.zeros(cur, dtype=np.float32)
    for n, off in zip(names, offsets):
        payload[off:off + tensors[n].size] = tensors[n].reshape(-1)
    return bytes(head) + payload.tobytes()


def deserialize(blob: bytes) -> "OrderedDict[str, np.ndarray]":
    if blob[:8] != MAGIC:
        raise ValueError("not a DPDFNet-B200 weight blob")
    (n,) = struct.unpack_from("<q", blob, 8)
    rec = NAME_LEN + 16
    head = 16 + n * rec
    head += (-head) % 128
    payload = np.frombuffer(blob, dtype=np.float32, offset=head)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for i in range(n):
        base = 16 + i * rec
        name = blob[base:base + NAME_LEN].split(b"\0", 1)[0].decode()
        off, numel = struct.unpack_from("<qq", blob, base + NAME_LEN)
        out[name] = payload[off:off + numel]
    return out


def pack_checkpoint(spec: ModelSpec, sd: Mapping[str, np.ndarray]) -> bytes:
    return serialize(pack_tensors(spec, sd))


def load_checkpoint_file(path) -> Dict[str, np.ndarray]:
    """Load a reference ``.pth`` (plain ``state_dict`` or ``{'state_dict': ...}``) as numpy."""
    import os
    import torch
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:
        # The reference's inference package never unpickles; only its export script falls back to a full pickle load
        # (export_dpdfnet_to_onnx.py:103-106).  Arbitrary-code pickles are therefore an explicit opt-in here.
        if os.environ.get("DPDFNET_B200_ALLOW_PICKLE") != "1":
            raise ValueError(f"{path} is not a plain tensor state_dict (safe load failed: {exc}); set "
                             "DPDFNET_B200_ALLOW_PICKLE=1 to unpickle a trusted checkpoint") from exc
        obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "state_dict" in obj and not any(k.startswith("enc.") for k in obj):
        obj = obj["state_dict"]
    return {k: v.detach().cpu().numpy() for k, v in obj.items() if hasattr(v, "detach")}
