"""CPU-side checks: weight packing, blob format, C-ABI surface (no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from dpdfnet_b200 import weights
from dpdfnet_b200.spec import MODEL_SPECS, get_spec

ROOT = Path(__file__).resolve().parents[1]


def test_blob_roundtrip():
    spec = get_spec("dpdfnet2")
    t = weights.pack_tensors(spec, weights.random_checkpoint(spec, 1))
    blob = weights.serialize(t)
    back = weights.deserialize(blob)
    assert list(back) == list(t)
    for k in t:
        assert np.array_equal(back[k], t[k].reshape(-1)), k


def test_random_checkpoint_is_deterministic_and_bn_randomised():
    spec = get_spec("dpdfnet4")
    a, b = weights.random_checkpoint(spec, 3), weights.random_checkpoint(spec, 3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["enc.erb_conv0.2.running_var"].std() > 0.1 and abs(a["enc.erb_conv0.2.running_mean"]).max() > 0.05


def test_pack_rejects_bad_checkpoints():
    spec = get_spec("dpdfnet2")
    ck = weights.random_checkpoint(spec, 0)
    bad = dict(ck)
    bad.pop("enc.df_conv1.1.weight")
    with pytest.raises(KeyError):
        weights.pack_tensors(spec, bad)
    bad = dict(ck)
    bad["enc.df_conv1.1.weight"] = bad["enc.df_conv1.1.weight"][:32]
    with pytest.raises(ValueError):
        weights.pack_tensors(spec, bad)
    with pytest.raises(ValueError):
        weights.deserialize(b"nonsense" * 10)


def test_dft_bases_reconstruct():
    """Vorbis COLA (package/tests/test_package_behaviors.py:709-716) through the packed bases:
    analysis -> synthesis -> overlap-add of two frames reproduces the middle hop."""
    for name in ("dpdfnet2", "dpdfnet2_48khz_hr"):
        spec = get_spec(name)
        b = weights.dft_bases(spec)
        fwd, inv = b["const.dft_fwd"], b["const.dft_inv"]
        x = np.random.default_rng(0).standard_normal(3 * spec.hop)
        frames = []
        for t in range(2):
            fr = x[t * spec.hop:t * spec.hop + spec.win]
            X = np.stack([fr @ fwd[:, :, 0], fr @ fwd[:, :, 1]], -1)
            frames.append(X[:, 0] @ inv[:, :, 0] + X[:, 1] @ inv[:, :, 1])
        mid = frames[0][spec.hop:] + frames[1][:spec.hop]
        assert np.abs(mid - x[spec.hop:2 * spec.hop]).max() < 1e-9


def test_c_abi_exports_every_declared_symbol():
    from dpdfnet_b200 import engine
    from dpdfnet_b200.build import build_library
    lib = engine.load_library(build_library())
    header = (ROOT / "include" / "dpdfnet_b200.h").read_text()
    declared = set(re.findall(r"\b(dpdf_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert declared == set(engine.C_API), declared ^ set(engine.C_API)
    assert lib.dpdf_version().decode().startswith("dpdfnet_b200")
    assert ctypes.sizeof(engine._Spec) == 4 * (8 + 4 + 3 + 3 + 32 + 1)


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dpdfnet_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        Engine("dpdfnet2", max_streams=1)


def test_c_spec_matches_python_spec():
    from dpdfnet_b200.engine import c_spec
    for name, spec in MODEL_SPECS.items():
        s = c_spec(spec)
        assert s.state_size == spec.state_size and sum(s.erb_widths) == spec.freq_bins
        assert list(s.fe) == list(spec.fe)


def test_fp16_split_operand_images():
    """Host side of the tcgen05 path: hi + lo reproduces FP32 weights to ~2^-22, the K-major SWIZZLE_NONE image puts
    element (n, k) at byte (n>>3)*SBO + (k>>3)*128 + (n&7)*16 + (k&7)*2, and out-of-range weights are refused."""
    import pytest
    from dpdfnet_b200.weights import fp16_split, umma_kmajor16, umma_operand16
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((192, 64)) * np.exp(rng.uniform(-8, 2, (192, 64)))).astype(np.float32)
    hi, lo = fp16_split(w)
    assert hi.dtype == np.float16 and lo.dtype == np.float16
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - w)
    assert np.all(err <= np.maximum(np.abs(w) * 2.0 ** -21, 2.0 ** -24))       # FP16-subnormal floor for tiny weights
    img = umma_kmajor16(hi).view(np.uint16)
    sbo = (64 // 8) * 128
    for n, k in [(0, 0), (7, 7), (8, 0), (13, 42), (191, 63), (100, 8)]:
        off = (n >> 3) * sbo + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2
        assert img[off // 2] == hi.view(np.uint16)[n, k]
    op = umma_operand16(w)
    assert op.dtype == np.float32 and op.size == 192 * 64                       # hi | lo halves, as raw f32 words
    with pytest.raises(ValueError):
        fp16_split(np.array([[7.0e4]], np.float32))


def test_packed_blob_carries_tensor_core_images():
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_tensors, random_checkpoint, LOG2E
    spec = get_spec("dpdfnet2")
    t = pack_tensors(spec, random_checkpoint(spec, 3))
    for br in ("erb", "df"):
        for i in range(spec.n_blocks):
            q = f"enc.dprnn_{br}.{i}"
            assert t[f"{q}.tc.intra"].size == 2 * 4 * 192 * 64 // 2
            assert t[f"{q}.tc.gates"].size == 6 * 64 * 64 and t[f"{q}.tc.fc_w"].size == 2 * 64 * 64
            b, tb = t[f"{q}.intra.bias"].reshape(2, 4, 64), t[f"{q}.tc.intra_bias"].reshape(2, 4, 64)
            assert np.allclose(tb[:, :2], -LOG2E * b[:, :2], rtol=1e-6) and np.allclose(tb[:, 2:], 2 * LOG2E * b[:, 2:], rtol=1e-6)


def test_fragment_form_images_permute_only_the_recurrent_k_axis():
    """tc.intra_f (k_dprnn_intra_tc.cu:intra_sweep_f): W_ih images identical to tc.intra; W_hh images hold, at K element
    k = 2 c + e of operand column c = 16 p + 4 cg + j, hidden unit 16 (2 p + e) + 4 cg + j - the two units one gate thread
    packs into that column - for both branches and both directions."""
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_tensors, random_checkpoint
    spec = get_spec("dpdfnet2_48khz_hr")
    t = pack_tensors(spec, random_checkpoint(spec, 4))
    W = 192 * 64                                            # halves per image

    def unimage(img16):                                     # K-major SWIZZLE_NONE image -> [192, 64]
        return img16.reshape(192 // 8, 64 // 8, 8, 8).transpose(0, 2, 1, 3).reshape(192, 64)

    for br in ("erb", "df"):
        for i in range(spec.n_blocks):
            a = t[f"enc.dprnn_{br}.{i}.tc.intra"].view(np.float16).reshape(2, 4, W)
            f = t[f"enc.dprnn_{br}.{i}.tc.intra_f"].view(np.float16).reshape(2, 4, W)
            assert np.array_equal(a[:, :2], f[:, :2])                             # W_ih hi | lo untouched
            for d in range(2):
                for im in (2, 3):                                                 # W_hh hi, lo
                    plain, perm = unimage(a[d, im]), unimage(f[d, im])
                    for c in range(32):
                        p, cg, j = c >> 4, (c >> 2) & 3, c & 3
                        for e in range(2):
                            assert np.array_equal(perm[:, 2 * c + e], plain[:, 16 * (2 * p + e) + 4 * cg + j])
