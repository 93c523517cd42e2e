"""CPU oracle for the DPDFNet per-frame hot path (numpy, float32, batched over streams).

TEST INFRASTRUCTURE ONLY.  Nothing under ``dpdfnet_b200/`` may import this module; it is the
checker for ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.

It restates, stage by stage, what the reference's streaming model computes for one hop
(``onnx_model/dpdfnet.py:748-806`` and ``onnx_model/dpdfnet_48khz_hr.py`` for 48 kHz), plus the
causal STFT / overlap-add of ``package/src/dpdfnet/stream.py:117-156``, in the engine's data
layout (activations ``[B, F, C]``, rings addressed by a per-stream frame counter) and from the
engine's packed weight tensors (``dpdfnet_b200.weights.pack_tensors``), so that BatchNorm
folding / weight packing are covered by the CPU parity check as well.

Parity pin: ``oracle/make_golden.py`` runs this oracle against the reference PyTorch modules
imported from ``/root/reference`` (streaming ``onnx_model`` per frame incl. the flat state
vector, and the offline ``model/dpdfnet.py`` on whole clips) and commits the vectors under
``tests/golden/``.  The reference's own tests hold no numerical golden for the network
(SURVEY.md section 4), so those generated vectors are the pin.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import numpy as np

from dpdfnet_b200.spec import (ALPHA_NORM, CONV_CH, DF_ORDER, GRU_DIM, NB_DF, ModelSpec)

FLAG_WARMUP = 1      # network not run: feature/c0/coef ring slots <- 0, recurrent states untouched
FLAG_ZERO_FEAT = 2   # normalised features forced to zero (offline look-ahead padding)
FLAG_PRIME = 4       # step_pcm only: store the hop as analysis history, nothing else
FLAG_ZERO_SPEC = 8   # step_pcm only: the analysed spectrum is taken as zero (offline DF look-ahead padding)

f32 = np.float32


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x, dtype=f32))).astype(f32)


def _layernorm(x, g, b, eps=1e-5):
    # torch.nn.LayerNorm over the last dim, biased variance (modules.py:92, layers.py:134)
    mu = x.mean(-1, keepdims=True, dtype=f32)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True, dtype=f32)
    return (xc / np.sqrt(var + f32(eps)) * g + b).astype(f32)


def gru_cell(x, h, wih, whh, bias):
    """torch.nn.GRUCell (layers.py:1211; gate order r,z,n).  bias = [4,H]: r, z, in, hn."""
    H = h.shape[-1]
    gi = x @ wih.T
    gh = h @ whh.T
    r = _sigmoid(gi[..., :H] + gh[..., :H] + bias[0])
    z = _sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H] + bias[1])
    n = np.tanh(gi[..., 2 * H:] + bias[2] + r * (gh[..., 2 * H:] + bias[3]), dtype=f32)
    return ((1.0 - z) * n + z * h).astype(f32)


def grouped_linear(x, w, b):
    """GroupedLinear (layers.py:1020-1046).  w: [G, O/G, I/G], b: [O]."""
    G, og, ig = w.shape
    xs = x.reshape(x.shape[0], G, ig)
    y = np.einsum("bgi,goi->bgo", xs, w, dtype=f32)
    return (y.reshape(x.shape[0], G * og) + b).astype(f32)


def sepconv(x, dw, pw, b, stride=1):
    """Depthwise 1x3 (freq pad 1, freq stride) [+ sub-pixel interleave] -> pointwise -> +b -> ReLU.

    x: [B, F, C]; dw: [S, 3, C]; Conv2dNormAct / SubPixelConv2dNormAct, layers.py:761-834, 895-973.
    """
    B, F, C = x.shape
    S = dw.shape[0]
    xp = np.zeros((B, F + 2, C), f32)
    xp[:, 1:F + 1] = x
    if S == 1:
        fo = (F - 1) // stride + 1
        idx = np.arange(fo) * stride
        t = xp[:, idx] * dw[0, 0] + xp[:, idx + 1] * dw[0, 1] + xp[:, idx + 2] * dw[0, 2]
    else:
        assert stride == 1
        t = np.zeros((B, F, S, C), f32)
        for j in range(S):
            t[:, :, j] = xp[:, 0:F] * dw[j, 0] + xp[:, 1:F + 1] * dw[j, 1] + xp[:, 2:F + 2] * dw[j, 2]
        t = t.reshape(B, F * S, C)          # out[f*S + j] = conv_j[f]  (layers.py:915)
    y = t.astype(f32) @ pw.T + b
    return np.maximum(y, 0).astype(f32)


class OracleEngine:
    def __init__(self, spec: ModelSpec, tensors: Mapping[str, np.ndarray], max_streams: int):
        self.spec = spec
        self.B = int(max_streams)
        self.w: Dict[str, np.ndarray] = {}
        shapes = self._tensor_shapes()
        for k, v in tensors.items():
            self.w[k] = np.asarray(v, f32).reshape(shapes[k]) if k in shapes else np.asarray(v, f32)
        widths = spec.erb_widths
        self.band_of_bin = np.repeat(np.arange(len(widths)), widths)
        self.band_start = np.concatenate([[0], np.cumsum(widths)[:-1]])
        self.band_inv_w = (1.0 / np.asarray(widths, np.float64)).astype(f32)
        self.dbg: Dict[str, np.ndarray] = {}
        self._alloc()
        self.reset()

    # ------------------------------------------------------------------
    def _tensor_shapes(self):
        sp, C, H = self.spec, CONV_CH, GRU_DIM
        F, N = sp.freq_bins, sp.win
        s = {"const.dft_fwd": (N, F, 2), "const.dft_inv": (F, N, 2), "const.mu0": (sp.fe_feat,), "const.s0": (NB_DF,),
             "enc.erb_conv0.w": (9, C), "enc.df_conv0.w": (9, C), "enc.df_conv0.pw": (C, C),
             "erb_dec.conv0_out.w": (3, C), "df_dec.df_convp.w": (10, 5, 32), "df_dec.df_convp.pw": (10, 10)}
        for n in ["enc.erb_conv1", "enc.erb_conv2", "enc.erb_conv3", "enc.df_conv1"]:
            s[n + ".dw"], s[n + ".pw"] = (1, 3, C), (C, C)
        for i, up in zip((3, 2, 1), sp.dec_up):
            s[f"erb_dec.convt{i}.dw"], s[f"erb_dec.convt{i}.pw"] = (up, 3, C), (C, C)
        for br in ("erb", "df"):
            for i in range(sp.n_blocks):
                q = f"enc.dprnn_{br}.{i}"
                s[q + ".intra.wih"] = s[q + ".intra.whh"] = (2, 3 * C, C)
                s[q + ".intra.bias"] = (2, 4, C)
                s[q + ".intra.fc_w"] = (C, 2 * C)
                s[q + ".inter.wih"] = s[q + ".inter.whh"] = (3 * C, C)
                s[q + ".inter.bias"] = (4, C)
                s[q + ".inter.fc_w"] = (C, C)
        gls = {"enc.df_fc_emb": (32, 16, 96), "enc.emb_gru.lin_in": (16, 16, 64), "enc.emb_gru.lin_out": (16, 32, 16),
               "erb_dec.emb_gru.lin_in": (16, 16, 32), "erb_dec.emb_gru.lin_out": (16, 32, 16),
               "df_dec.df_gru.lin_in": (8, 32, 64), "df_dec.df_skip": (16, 16, 32), "df_dec.df_out": (16, 60, 16)}
        if sp.hr48:
            k = C * sp.fe[3] // 32
            gls["enc.erb_fc_emb"] = (32, 16, k)
            gls["erb_dec.erb_fc_emb"] = (32, k, 16)
        for n, shp in gls.items():
            s[n + ".w"] = shp
        for n, layers in (("enc.emb_gru.gru", 1), ("erb_dec.emb_gru.gru", 2), ("df_dec.df_gru.gru", 2)):
            for l in range(layers):
                s[f"{n}.{l}.wih"] = s[f"{n}.{l}.whh"] = (3 * H, H)
                s[f"{n}.{l}.bias"] = (4, H)
        return s

    def _alloc(self):
        sp, B, C = self.spec, self.B, CONV_CH
        F, N = sp.freq_bins, sp.n_blocks
        z = lambda *shape: np.zeros((B,) + shape, f32)
        self.mu, self.s = z(sp.fe_feat), z(NB_DF)
        self.erb_ring, self.df_ring = z(3, sp.fe_feat), z(3, 2, NB_DF)
        self.inter_erb, self.inter_df = z(N, sp.fe[3], C), z(N, NB_DF // 2, C)
        self.h_enc, self.h_erb, self.h_df = z(GRU_DIM), z(2, GRU_DIM), z(2, GRU_DIM)
        self.c0_ring = z(DF_ORDER, NB_DF, C)
        self.mask_ring, self.coef_ring, self.dfspec_ring = z(3, F, 2), z(3, NB_DF, 2 * DF_ORDER), z(DF_ORDER, F, 2)
        self.in_hist, self.ola = z(sp.hop), z(sp.hop)
        self.pos = np.zeros(B, np.int64)

    def reset(self, slots=None):
        sl = slice(None) if slots is None else np.asarray(slots)
        for a in (self.erb_ring, self.df_ring, self.inter_erb, self.inter_df, self.h_enc, self.h_erb, self.h_df,
                  self.c0_ring, self.mask_ring, self.coef_ring, self.dfspec_ring, self.in_hist, self.ola):
            a[sl] = 0
        self.mu[sl] = self.w["const.mu0"]
        self.s[sl] = self.w["const.s0"]
        self.pos[sl] = 0

    # ----- ring helpers -----------------------------------------------
    @staticmethod
    def _ring_write(ring, slots, pos, value):
        L = ring.shape[1]
        ring[slots, pos % L] = value

    @staticmethod
    def _ring_logical(ring, slots, pos):
        """Frames oldest-first after this step's write (CyclicBuffer order, layers.py:99-103)."""
        L = ring.shape[1]
        idx = (pos[:, None] + 1 + np.arange(L)[None, :]) % L
        return ring[slots[:, None], idx]

    # ----- reference flat state (onnx_model/dpdfnet.py:737-746) ----------
    def export_state(self, slot: int) -> np.ndarray:
        p = int(self.pos[slot])
        log = lambda ring: np.stack([ring[slot, (p + k) % ring.shape[1]] for k in range(ring.shape[1])], 0)
        N = self.spec.n_blocks
        parts = [self.mu[slot], self.s[slot], log(self.erb_ring)]
        parts += [self.inter_erb[slot, i] for i in range(N)]
        parts += [log(self.df_ring)]
        parts += [self.inter_df[slot, i] for i in range(N)]
        parts += [self.h_enc[slot], self.h_erb[slot], self.h_df[slot]]
        parts += [log(self.c0_ring).transpose(0, 2, 1)]                          # [5, C, F]
        parts += [log(self.mask_ring)]
        parts += [log(self.coef_ring).reshape(3, NB_DF, DF_ORDER, 2).transpose(0, 2, 1, 3)]   # [3, 5, F, 2]
        parts += [log(self.dfspec_ring)]
        return np.concatenate([np.asarray(x, f32).reshape(-1) for x in parts])

    def import_state(self, slot: int, flat: np.ndarray):
        flat = np.asarray(flat, f32).reshape(-1)
        if flat.size != self.spec.state_size:
            raise ValueError(f"state size mismatch: expected {self.spec.state_size}, got {flat.size}")
        off = 0
        seg = {}
        for name, shape in self.spec.state_segments():
            n = int(np.prod(shape))
            seg[name] = flat[off:off + n].reshape(shape)
            off += n
        N = self.spec.n_blocks
        self.pos[slot] = 0        # logical frame k lives in physical slot k
        self.mu[slot], self.s[slot] = seg["erb_norm.mu"], seg["spec_norm.s"]
        self.erb_ring[slot], self.df_ring[slot] = seg["enc.erb_conv0.ring"], seg["enc.df_conv0.ring"]
        for i in range(N):
            self.inter_erb[slot, i] = seg[f"enc.dprnn_erb.{i}.h"]
            self.inter_df[slot, i] = seg[f"enc.dprnn_df.{i}.h"]
        self.h_enc[slot], self.h_erb[slot], self.h_df[slot] = seg["enc.emb_gru.h"], seg["erb_dec.emb_gru.h"], seg["df_dec.df_gru.h"]
        self.c0_ring[slot] = seg["df_dec.c0.ring"].transpose(0, 2, 1)
        self.mask_ring[slot] = seg["mask.ring"]
        self.coef_ring[slot] = seg["df_op.coef.ring"].transpose(0, 2, 1, 3).reshape(3, NB_DF, 2 * DF_ORDER)
        self.dfspec_ring[slot] = seg["df_op.spec.ring"]

    # ----- stages --------------------------------------------------------
    def _features(self, X, sl):
        """a2-a4: ERB / magnitude features + running norms (onnx_model/dpdfnet.py:814-852)."""
        sp = self.spec
        a, one_m_a = f32(ALPHA_NORM), f32(1 - ALPHA_NORM)
        re, im = X[..., 0], X[..., 1]
        pw = (re * re + im * im).astype(f32)
        if sp.hr48:
            feat = 10 * np.log10(np.sqrt(pw) + f32(1e-10), dtype=f32)
        else:
            band = np.add.reduceat(pw * self.band_inv_w[self.band_of_bin], self.band_start, axis=1).astype(f32)
            feat = 10 * np.log10(band + f32(1e-10), dtype=f32)
        mu = (a * self.mu[sl] + one_m_a * feat).astype(f32)
        fe = ((feat - mu) / f32(40.0)).astype(f32)
        mag = np.sqrt(pw[:, :NB_DF])
        s = (a * self.s[sl] + one_m_a * mag).astype(f32)
        den = np.sqrt(s + f32(1e-12))
        fs = np.stack([re[:, :NB_DF] / den, im[:, :NB_DF] / den], 1).astype(f32)     # [B, 2, 96]
        self.mu[sl], self.s[sl] = mu, s
        return fe, fs

    def _encoder_convs(self, erb_log, df_log):
        """a5.  erb_log [B,3,fe_feat], df_log [B,3,2,96] (oldest first)."""
        sp, w, C = self.spec, self.w, CONV_CH
        B = erb_log.shape[0]
        fe0 = sp.fe[0]
        xp = np.zeros((B, 3, fe0 + 2), f32)
        xp[:, :, 1:fe0 + 1] = erb_log[:, :, :fe0]
        e0 = np.zeros((B, fe0, C), f32)
        for kt in range(3):
            for kf in range(3):
                e0 += xp[:, kt, kf:kf + fe0, None] * w["enc.erb_conv0.w"][kt * 3 + kf]
        e0 = np.maximum(e0 + w["enc.erb_conv0.b"], 0).astype(f32)
        s1, s2, s3 = sp.erb_strides
        e1 = sepconv(e0, w["enc.erb_conv1.dw"], w["enc.erb_conv1.pw"], w["enc.erb_conv1.b"], s1)
        e2 = sepconv(e1, w["enc.erb_conv2.dw"], w["enc.erb_conv2.pw"], w["enc.erb_conv2.b"], s2)
        e3 = sepconv(e2, w["enc.erb_conv3.dw"], w["enc.erb_conv3.pw"], w["enc.erb_conv3.b"], s3)
        dp = np.zeros((B, 3, 2, NB_DF + 2), f32)
        dp[..., 1:NB_DF + 1] = df_log
        t0 = np.zeros((B, NB_DF, C), f32)
        gw = w["enc.df_conv0.w"]
        for kt in range(3):
            for kf in range(3):
                t0[:, :, :C // 2] += dp[:, kt, 0, kf:kf + NB_DF, None] * gw[kt * 3 + kf, :C // 2]
                t0[:, :, C // 2:] += dp[:, kt, 1, kf:kf + NB_DF, None] * gw[kt * 3 + kf, C // 2:]
        c0 = np.maximum(t0 @ w["enc.df_conv0.pw"].T + w["enc.df_conv0.b"], 0).astype(f32)
        c1 = sepconv(c0, w["enc.df_conv1.dw"], w["enc.df_conv1.pw"], w["enc.df_conv1.b"], 2)
        return e0, e1, e2, e3, c0, c1

    def _dprnn_block(self, x, hstate, q):
        """a6/a7: one DPRNNBlock (layers.py:159-196).  x [B,F',C]; hstate [B,F',C] inter state."""
        w = self.w
        B, Fp, C = x.shape
        hcat = np.zeros((B, Fp, 2 * C), f32)
        for d in range(2):
            h = np.zeros((B, C), f32)
            order = range(Fp) if d == 0 else range(Fp - 1, -1, -1)
            for f in order:
                h = gru_cell(x[:, f], h, w[q + ".intra.wih"][d], w[q + ".intra.whh"][d], w[q + ".intra.bias"][d])
                hcat[:, f, d * C:(d + 1) * C] = h
        y = _layernorm(hcat @ w[q + ".intra.fc_w"].T + w[q + ".intra.fc_b"], w[q + ".intra.ln_g"], w[q + ".intra.ln_b"]) + x
        y = y.astype(f32)
        hn = gru_cell(y.reshape(B * Fp, C), hstate.reshape(B * Fp, C), w[q + ".inter.wih"], w[q + ".inter.whh"],
                      w[q + ".inter.bias"]).reshape(B, Fp, C)
        z = _layernorm(hn @ w[q + ".inter.fc_w"].T + w[q + ".inter.fc_b"], w[q + ".inter.ln_g"], w[q + ".inter.ln_b"]) + y
        return z.astype(f32), hn, hcat

    def _network(self, erb_log, df_log, sl, commit):
        """a5-a10 for the streams in ``sl``; ``commit`` [B] bool gates recurrent-state writes."""
        sp, w, C = self.spec, self.w, CONV_CH
        dbg = self.dbg
        B = erb_log.shape[0]
        e0, e1, e2, e3, c0, c1 = self._encoder_convs(erb_log, df_log)
        dbg.update(e0=e0, e1=e1, e2=e2, e3=e3, c0=c0, c1=c1)
        cm = commit[:, None, None]
        xe, xd = e3, c1
        for i in range(sp.n_blocks):
            xe, hn, hcat = self._dprnn_block(xe, self.inter_erb[sl, i], f"enc.dprnn_erb.{i}")
            self.inter_erb[sl, i] = np.where(cm, hn, self.inter_erb[sl, i])
            dbg[f"erb_hcat{i}"], dbg[f"xe{i}"] = hcat, xe
            xd, hn, hcat = self._dprnn_block(xd, self.inter_df[sl, i], f"enc.dprnn_df.{i}")
            self.inter_df[sl, i] = np.where(cm, hn, self.inter_df[sl, i])
            dbg[f"df_hcat{i}"], dbg[f"xd{i}"] = hcat, xd
        relu = lambda v: np.maximum(v, 0).astype(f32)
        # a8: embedding + encoder GRU (onnx_model/dpdfnet.py:233-241)
        emb_e = xe.reshape(B, -1)
        if sp.hr48:
            emb_e = relu(grouped_linear(emb_e, w["enc.erb_fc_emb.w"], w["enc.erb_fc_emb.b"]))
        cemb = relu(grouped_linear(xd.reshape(B, -1), w["enc.df_fc_emb.w"], w["enc.df_fc_emb.b"]))
        emb_in = np.concatenate([emb_e, cemb], -1)
        x = relu(grouped_linear(emb_in, w["enc.emb_gru.lin_in.w"], w["enc.emb_gru.lin_in.b"]))
        h = gru_cell(x, self.h_enc[sl], w["enc.emb_gru.gru.0.wih"], w["enc.emb_gru.gru.0.whh"], w["enc.emb_gru.gru.0.bias"])
        self.h_enc[sl] = np.where(commit[:, None], h, self.h_enc[sl])
        emb = relu(grouped_linear(h, w["enc.emb_gru.lin_out.w"], w["enc.emb_gru.lin_out.b"]))
        dbg.update(cemb=cemb, emb=emb)
        # a9: ERB decoder (onnx_model/dpdfnet.py:343-368)
        x = relu(grouped_linear(emb, w["erb_dec.emb_gru.lin_in.w"], w["erb_dec.emb_gru.lin_in.b"]))
        for l in range(2):
            x = gru_cell(x, self.h_erb[sl, l], w[f"erb_dec.emb_gru.gru.{l}.wih"], w[f"erb_dec.emb_gru.gru.{l}.whh"],
                         w[f"erb_dec.emb_gru.gru.{l}.bias"])
            self.h_erb[sl, l] = np.where(commit[:, None], x, self.h_erb[sl, l])
        ed = relu(grouped_linear(x, w["erb_dec.emb_gru.lin_out.w"], w["erb_dec.emb_gru.lin_out.b"]))
        if sp.hr48:
            ed = relu(grouped_linear(ed, w["erb_dec.erb_fc_emb.w"], w["erb_dec.erb_fc_emb.b"]))
        ed = ed.reshape(B, sp.fe[3], C)
        path = lambda e, n: relu(e * w[f"erb_dec.{n}.a"] + w[f"erb_dec.{n}.b"])
        d3 = sepconv(path(e3, "conv3p") + ed, w["erb_dec.convt3.dw"], w["erb_dec.convt3.pw"], w["erb_dec.convt3.b"])
        d2 = sepconv(path(e2, "conv2p") + d3, w["erb_dec.convt2.dw"], w["erb_dec.convt2.pw"], w["erb_dec.convt2.b"])
        d1 = sepconv(path(e1, "conv1p") + d2, w["erb_dec.convt1.dw"], w["erb_dec.convt1.pw"], w["erb_dec.convt1.b"])
        u = (path(e0, "conv0p") + d1).astype(f32)
        fe0 = sp.fe[0]
        up = np.zeros((B, fe0 + 2, C), f32)
        up[:, 1:fe0 + 1] = u
        wo = w["erb_dec.conv0_out.w"]
        mlin = (up[:, 0:fe0] * wo[0] + up[:, 1:fe0 + 1] * wo[1] + up[:, 2:fe0 + 2] * wo[2]).sum(-1, dtype=f32)
        m = _sigmoid(mlin + w["erb_dec.conv0_out.b"][0])
        dbg.update(ed=ed, d3=d3, d2=d2, d1=d1, m=m)
        # a10: DF decoder (onnx_model/dpdfnet.py:486-519)
        x = relu(grouped_linear(emb, w["df_dec.df_gru.lin_in.w"], w["df_dec.df_gru.lin_in.b"]))
        for l in range(2):
            x = gru_cell(x, self.h_df[sl, l], w[f"df_dec.df_gru.gru.{l}.wih"], w[f"df_dec.df_gru.gru.{l}.whh"],
                         w[f"df_dec.df_gru.gru.{l}.bias"])
            self.h_df[sl, l] = np.where(commit[:, None], x, self.h_df[sl, l])
        c = (x + grouped_linear(emb, w["df_dec.df_skip.w"], w["df_dec.df_skip.b"])).astype(f32)
        co = np.tanh(grouped_linear(c, w["df_dec.df_out.w"], w["df_dec.df_out.b"]), dtype=f32).reshape(B, NB_DF, 2 * DF_ORDER)
        dbg["co"] = co
        return m, co, c0

    def _df_pathway(self, c0_log):
        """df_convp on the 5-frame c0 ring: grouped (2 x 32->5, 5x1) + pointwise 10x10 + BN + ReLU."""
        w = self.w
        gw = w["df_dec.df_convp.w"]                                      # [10, 5, 32]
        B = c0_log.shape[0]
        t = np.zeros((B, NB_DF, 10), f32)
        for o in range(10):
            g = o // 5
            t[:, :, o] = np.einsum("btfc,tc->bf", c0_log[:, :, :, g * 32:(g + 1) * 32], gw[o], dtype=f32)
        return np.maximum(t @ w["df_dec.df_convp.pw"].T + w["df_dec.df_convp.b"], 0).astype(f32)

    # ----- one hop ---------------------------------------------------------
    def _step_core(self, X, slots, flags):
        """X [B,F,2] already scaled by wnorm.  Returns the enhanced, still scaled spectrum."""
        sp = self.spec
        B = X.shape[0]
        pos = self.pos[slots].copy()
        warm = (flags & FLAG_WARMUP) != 0
        zf = (flags & FLAG_ZERO_FEAT) != 0
        self.dbg = {}
        fe, fs = self._features(X, slots)
        self.dbg.update(spec=X.copy(), feat_erb=fe.copy(), feat_spec=fs.copy())
        kill = (warm | zf)
        fe = np.where(kill[:, None], f32(0), fe)
        fs = np.where(kill[:, None, None], f32(0), fs)
        self._ring_write(self.erb_ring, slots, pos, fe)
        self._ring_write(self.df_ring, slots, pos, fs)
        m, co, c0 = self._network(self._ring_logical(self.erb_ring, slots, pos),
                                  self._ring_logical(self.df_ring, slots, pos), slots, ~warm)
        c0 = np.where(warm[:, None, None], f32(0), c0)
        self._ring_write(self.c0_ring, slots, pos, c0)
        coefs = (co + self._df_pathway(self._ring_logical(self.c0_ring, slots, pos))).astype(f32)
        coefs = np.where(warm[:, None, None], f32(0), coefs)
        m = np.where(warm[:, None], f32(0), m)
        self._ring_write(self.coef_ring, slots, pos, coefs)
        self.dbg["coefs"] = coefs
        # a11: mask with a 2-frame spectrum delay (layers.py:414-445 / dpdfnet_48khz_hr.py:55-69)
        self._ring_write(self.mask_ring, slots, pos, X)
        Xd = self._ring_logical(self.mask_ring, slots, pos)[:, 0]
        if sp.hr48:
            gain = np.concatenate([m, m[:, -2:-1]], 1)        # reflect pad: m[480] = m[478]
        else:
            gain = m[:, self.band_of_bin]
        S = (Xd * gain[..., None]).astype(f32)
        self._ring_write(self.dfspec_ring, slots, pos, S)
        # a12: deep filter (onnx_model/multiframe.py:140-154, 200-232)
        Sl = self._ring_logical(self.dfspec_ring, slots, pos)             # [B,5,F,2]
        cd = self._ring_logical(self.coef_ring, slots, pos)[:, 0].reshape(B, NB_DF, DF_ORDER, 2)
        sr, si = Sl[:, :, :NB_DF, 0], Sl[:, :, :NB_DF, 1]
        cr, ci = cd[..., 0].transpose(0, 2, 1), cd[..., 1].transpose(0, 2, 1)
        Y = Sl[:, 2].copy()
        Y[:, :NB_DF, 0] = (sr * cr).sum(1, dtype=f32) - (si * ci).sum(1, dtype=f32)
        Y[:, :NB_DF, 1] = (sr * ci).sum(1, dtype=f32) + (si * cr).sum(1, dtype=f32)
        self.pos[slots] = pos + 1
        self.dbg.update(gain=gain, masked=S, spec_out=Y.copy())
        return Y

    def step_spec(self, spec_in, slots=None, flags=None):
        """ONNX-shaped entry: un-normalised spectrum in / out (export_dpdfnet_to_onnx.py:21-25)."""
        spec_in = np.asarray(spec_in, f32)
        B = spec_in.shape[0]
        slots = np.arange(B) if slots is None else np.asarray(slots)
        flags = np.zeros(B, np.int64) if flags is None else np.asarray(flags, np.int64)
        wn = f32(self.spec.wnorm)
        Y = self._step_core((spec_in * wn).astype(f32), slots, flags)
        return (Y * f32(1.0 / self.spec.wnorm)).astype(f32)

    def step_pcm(self, pcm, slots=None, flags=None):
        """Causal streaming entry: hop samples in, hop samples out (stream.py:117-156)."""
        pcm = np.asarray(pcm, f32)
        B, hop = pcm.shape
        assert hop == self.spec.hop
        slots = np.arange(B) if slots is None else np.asarray(slots)
        flags = np.zeros(B, np.int64) if flags is None else np.asarray(flags, np.int64)
        prime = (flags & FLAG_PRIME) != 0
        out = np.zeros((B, hop), f32)
        run = np.nonzero(~prime)[0]
        if run.size:
            sl = slots[run]
            frame = np.concatenate([self.in_hist[sl], pcm[run]], 1)
            fwd = self.w["const.dft_fwd"]
            X = np.stack([frame @ fwd[:, :, 0], frame @ fwd[:, :, 1]], -1).astype(f32)
            X[(flags[run] & FLAG_ZERO_SPEC) != 0] = 0
            Y = self._step_core(X, sl, flags[run])
            inv = self.w["const.dft_inv"]
            t = (Y[..., 0] @ inv[:, :, 0] + Y[..., 1] @ inv[:, :, 1]).astype(f32)
            out[run] = self.ola[sl] + t[:, :hop]
            self.ola[sl] = t[:, hop:]
        self.in_hist[slots] = pcm
        return out


def make_oracle(spec: ModelSpec, checkpoint: Mapping[str, np.ndarray], max_streams: int) -> OracleEngine:
    from dpdfnet_b200.weights import pack_tensors
    return OracleEngine(spec, pack_tensors(spec, checkpoint), max_streams)


def offline_exact(eng: OracleEngine, wave: np.ndarray) -> np.ndarray:
    """Reproduce ``model/dpdfnet.py:DPDFNet.forward`` (offline, whole clip) with the streaming
    recurrence, using the offline-exact schedule of SURVEY.md section 7.  wave [B, n] -> [B, hop*(T-1)].
    """
    sp = eng.spec
    wave = np.asarray(wave, f32)
    B, n = wave.shape
    hop = sp.hop
    T = 1 + n // hop
    # torch.stft(center=True, pad_mode="reflect") framing: pad n_fft//2 = hop on both sides
    padded = np.concatenate([wave[:, hop:0:-1], wave, wave[:, -2:-hop - 2:-1]], 1)
    eng.reset(np.arange(B))
    z = np.zeros(B, np.int64)
    eng.step_pcm(padded[:, :hop], flags=z + FLAG_PRIME)
    outs = []
    for t in range(T + 4):
        if t < T:
            chunk = padded[:, (t + 1) * hop:(t + 2) * hop]
            fl = z + (FLAG_WARMUP if t < 2 else 0)
        else:
            chunk = np.zeros((B, hop), f32)
            # look-ahead frames are zero *spectra* (model/multiframe.py:74) with zero features (model/dpdfnet.py:463)
            fl = z + FLAG_ZERO_SPEC + (FLAG_ZERO_FEAT if t < T + 2 else 0)
        outs.append(eng.step_pcm(chunk, flags=fl))
    # step tau emits OLA hop tau of the output frames tau-4; keep hops 1..T-1 of the aligned signal
    return np.concatenate(outs[4 + 1:4 + T], 1)
