"""Parity of the CUDA engine (through the C ABI) against the CPU oracle, the golden vectors
produced by the reference PyTorch modules, and size-independent properties at bench sizes.

Tolerances: the north_star bar is 1e-4 max-abs on the enhanced waveform; everything here is
fp32 with different summation orders only, so per-frame spectra are held to 2e-4 in
un-normalised units (|X| ~ 30, i.e. ~7e-6 relative) and waveforms to 1e-4 (measured ~1e-6).
"""
import numpy as np
import pytest

from dpdfnet_b200.spec import get_spec
from dpdfnet_b200.weights import pack_tensors, random_checkpoint

pytestmark = pytest.mark.gpu

SPEC_TOL = 2e-4
WAVE_TOL = 1e-4
STATE_TOL = 1e-4


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    return torch


def _engine(name, seed, B, **kw):
    from dpdfnet_b200.engine import Engine
    spec = get_spec(name)
    return Engine(spec, random_checkpoint(spec, seed), max_streams=B, **kw)


def _oracle(name, seed, B):
    from oracle.oracle_np import OracleEngine
    spec = get_spec(name)
    return OracleEngine(spec, pack_tensors(spec, random_checkpoint(spec, seed)), B)


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("name", ["dpdfnet2", "dpdfnet4", "dpdfnet2_48khz_hr"])
def test_stream_golden(torch_cuda, golden_dir, name, tc):
    """Per-frame ONNX-shaped call with HOST buffers vs the reference streaming model's outputs (both kernel arms)."""
    g = np.load(golden_dir / f"stream_{name}.npz")
    eng = _engine(name, int(g["seed"]), 2)
    for opt in ("intra_tc", "post_tc", "sep_tc", "gru_tc", "dft_tc"):
        eng.set_option(opt, tc)
    for t in range(g["spec_in"].shape[0]):
        y = eng.step_spec_host(g["spec_in"][t][None], slot_ids=[1])
        assert np.isfinite(y).all()
        assert np.abs(y[0] - g["spec_out"][t]).max() < SPEC_TOL, t
    st = eng.state_export(1)
    assert st.shape == g["state"].shape
    assert np.abs(st - g["state"]).max() < STATE_TOL
    # untouched slot still holds the initial state
    init = eng.state_export(0)
    assert np.count_nonzero(init[eng.spec.fe_feat + 96:]) == 0


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("name", ["dpdfnet2", "dpdfnet4", "dpdfnet2_48khz_hr"])
def test_offline_golden(torch_cuda, golden_dir, name, tc):
    """BASELINE configs[0]-style check: whole clip vs model/dpdfnet.py output, 1e-4 max-abs -- on the small-batch
    FFMA2 kernels (tc=0) and with every tensor-core kernel forced on (tc=1: the WARMUP / ZERO_FEAT / ZERO_SPEC
    schedule flags go through the tcgen05 intra-GRU, post, separable-conv and GRU(256) kernels)."""
    from dpdfnet_b200.offline import enhance_offline_exact
    g = np.load(golden_dir / f"offline_{name}.npz")
    eng = _engine(name, int(g["seed"]), g["wave_in"].shape[0])
    for opt in ("intra_tc", "post_tc", "sep_tc", "gru_tc", "dft_tc"):
        eng.set_option(opt, tc)
    out = enhance_offline_exact(eng, g["wave_in"])
    assert out.shape == g["wave_out"].shape
    assert np.abs(out - g["wave_out"]).max() < WAVE_TOL


@pytest.mark.parametrize("intra_tc", [0, 1])
@pytest.mark.parametrize("name,B", [("dpdfnet2", 5), ("dpdfnet8", 3), ("baseline", 2), ("dpdfnet8_48khz_hr", 2)])
def test_stages_vs_oracle(torch_cuda, name, B, intra_tc):
    """Every intermediate tensor of a hop against the oracle, ragged batch + slot indirection; the intra-frame
    GRU on the FFMA2 kernel (intra_tc=0) and on the tcgen05 FP16-split kernel (intra_tc=1)."""
    eng = _engine(name, 11, B + 3)
    eng.set_option("graph", 0)
    eng.set_option("intra_tc", intra_tc)
    eng.set_option("sep_tc", intra_tc)          # the separable convs and the GRU(256) cells on the same arm: FFMA2 / tcgen05
    eng.set_option("gru_tc", intra_tc)
    eng.set_option("dfp_ps", intra_tc)          # df pathway conv: 5-frame ring form (default) / pending-partial-sum form
    ora = _oracle(name, 11, B + 3)
    spec = eng.spec
    rng = np.random.default_rng(2)
    slots = (np.arange(B)[::-1] + 2).astype(np.int32)
    N = spec.n_blocks
    for t in range(4):
        X = (rng.standard_normal((B, spec.freq_bins, 2)) * 20).astype(np.float32)
        y = eng.step_spec_host(X, slot_ids=slots)
        yo = ora.step_spec(X, slots=slots.astype(np.int64))
        assert np.abs(y - yo).max() < SPEC_TOL
        for stage, ref in (("e3", ora.dbg["e3"]), ("c0", ora.dbg["c0"]), ("emb", ora.dbg["emb"]),
                           ("m", ora.dbg["m"]), ("co", ora.dbg["coefs"] * 0 + ora.dbg["coefs"]),
                           ("xd", ora.dbg[f"xd{N - 1}"] if N else ora.dbg["c1"])):
            if stage == "co":
                continue          # 'co' is pre-pathway; covered through the coefficient ring in the state
            got = eng.debug_tensor(stage, B)
            assert np.abs(got - np.asarray(ref).reshape(B, -1)).max() < 2e-4, (stage, t)
    for b in range(B):
        assert np.abs(eng.state_export(int(slots[b])) - ora.export_state(int(slots[b]))).max() < STATE_TOL


def test_pcm_path_vs_oracle_and_graph(torch_cuda):
    """Fused DFT / OLA path: graph replay over T hops == T single steps == oracle."""
    torch = torch_cuda
    name, B, T = "dpdfnet2", 6, 12
    eng = _engine(name, 5, B)
    ora = _oracle(name, 5, B)
    hop = eng.spec.hop
    rng = np.random.default_rng(8)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    ref = np.concatenate([ora.step_pcm(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    out_graph = eng.run_pcm_host(pcm)
    assert np.abs(out_graph - ref).max() < WAVE_TOL
    eng.reset()
    eng.set_option("graph", 0)
    out_steps = np.concatenate([eng.step_pcm_host(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.array_equal(out_graph, out_steps)
    # device-pointer entry with a strided view
    eng.reset()
    eng.set_option("graph", 1)
    x = torch.from_numpy(pcm).cuda()
    y = eng.run_pcm(x)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), out_graph)


@pytest.mark.parametrize("name,B", [("dpdfnet2", 130), ("dpdfnet2_48khz_hr", 129)])
def test_intra_tc_multi_tile(torch_cuda, name, B):
    """tcgen05 intra-GRU kernel with more than one 128-stream tile and a ragged last tile: waveform and the
    dual-path activations against the oracle, and bit-equal rows for equal inputs across tiles."""
    T = 5
    eng = _engine(name, 3, B)
    eng.set_option("intra_tc", 1)
    ora = _oracle(name, 3, B)
    hop = eng.spec.hop
    rng = np.random.default_rng(12)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    pcm[128:] = pcm[:B - 128]                      # rows of the second tile repeat rows of the first
    ref = np.concatenate([ora.step_pcm(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    out = eng.run_pcm_host(pcm)
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() < WAVE_TOL
    assert np.array_equal(out[128:], out[:B - 128])
    N = eng.spec.n_blocks
    got = eng.debug_tensor("xd", B)
    assert np.abs(got - np.asarray(ora.dbg[f"xd{N - 1}"]).reshape(B, -1)).max() < 2e-4
    ffma = _engine(name, 3, B)
    ffma.set_option("intra_tc", 0)
    assert np.abs(ffma.run_pcm_host(pcm) - out).max() < 1e-5


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("name,B", [("dpdfnet2", 200), ("dpdfnet4", 330), ("dpdfnet2_48khz_hr", 70)])
def test_intra_tc_row_duplication(torch_cuda, name, B, overlap):
    """tcgen05 intra-GRU kernel with 128 / D streams per CTA (D rows of the MMA tile per stream, the gate math split
    over the D partner threads): D = 1, 2, 4 run the same arithmetic, so they must agree bit for bit, on ragged last
    tiles and both with the post kernel overlapped with the sweep (per-CTA progress counters, D per 128-stream post
    tile) and after it; the auto choice is one of them; all against the oracle."""
    T = 4
    hop = get_spec(name).hop
    rng = np.random.default_rng(31)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    outs = {}
    for D in (1, 2, 4, 0, "2sr", "4sr", "4f", "4fe"):
        eng = _engine(name, 4, B)
        eng.set_option("intra_tc", 1)
        eng.set_option("overlap", overlap)
        eng.set_option("lanes", 1)
        # split rows (intra_sr): the D rows of a stream carry the hi | lo halves of the operands instead of copies, two
        # MMA passes instead of three and the partner rows' accumulators are added in registers - the same products plus
        # the lo * lo term, in another order: equal to rounding, not bit for bit
        # fragment form (intra_frag, 32 streams per CTA): two rows per stream (hi | lo), .16x128b TMEM fragments hand each
        # of a stream's four gate threads its own column of both rows, K axis of W_hh permuted on the host
        if D != 0:                                          # 0: the engine's own choice of form
            eng.set_option("intra_sr", 1 if D in ("2sr", "4sr") else 0)
            eng.set_option("intra_frag", 1 if D in ("4f", "4fe") else 0)
            eng.set_option("intra_frag_erb", 1 if D == "4fe" else 0)        # erb sweep in fragment form too (48 kHz models only)
        eng.set_option("intra_dup", int(D[0]) if isinstance(D, str) else D)
        outs[D] = (eng.run_pcm_host(pcm), eng.debug_tensor("xd", B), eng.state_export(B - 1))
        eng.close()
    for D in (2, 4):
        for a, b in zip(outs[D], outs[1]):
            assert np.array_equal(a, b), D
    for a, b in zip(outs[0], outs[1]):                      # the auto choice may be a split-row form
        assert np.abs(a - b).max() < 2e-5
    for D in ("2sr", "4sr", "4f", "4fe"):
        assert np.abs(outs[D][0] - outs[1][0]).max() < 2e-6, D
        assert np.abs(outs[D][1] - outs[1][1]).max() < 2e-5, D
        assert np.abs(outs[D][2] - outs[1][2]).max() < 2e-5, D
    ora = _oracle(name, 4, B)
    ref = np.concatenate([ora.step_pcm(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    N = get_spec(name).n_blocks
    for D in (4, "2sr", "4sr", "4f", "4fe"):
        assert np.abs(outs[D][0] - ref).max() < WAVE_TOL, D
        assert np.abs(outs[D][1] - np.asarray(ora.dbg[f"xd{N - 1}"]).reshape(B, -1)).max() < 2e-4, D
    with pytest.raises(ValueError):
        _engine(name, 4, 2).set_option("intra_dup", 3)


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("name,B", [("dpdfnet2", 130), ("dpdfnet4", 300), ("dpdfnet2_48khz_hr", 67)])
def test_post_kernel_as_cta_pairs_and_dual_tiles(torch_cuda, name, B, overlap):
    """k_dprnn_post_tc<PAIR>: clusters of two CTAs issue every product as one tcgen05.mma.cta_group::2 of M = 256 over the
    pair's two tiles, each CTA holding half of every weight slab.  Same arithmetic per row as the single-CTA kernel, so
    the results must be bit-identical - with odd tile counts per branch (a padding tile closes the pair), ragged last
    tiles, after the sweep and overlapped with it - and match the oracle."""
    T = 4
    hop = get_spec(name).hop
    rng = np.random.default_rng(41)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    outs = {}
    for form in (None, "post_pair", "post_dual"):           # post_dual: two tiles per 1024-thread CTA sharing one 5-slab weight ring
        eng = _engine(name, 8, B)
        eng.set_option("intra_tc", 1)
        eng.set_option("overlap", overlap)
        eng.set_option("lanes", 1)
        if form:
            eng.set_option(form, 1)
        outs[form] = (eng.run_pcm_host(pcm), eng.debug_tensor("xd", B), eng.debug_tensor("xe", B), eng.state_export(B - 1))
        eng.close()
    for form in ("post_pair", "post_dual"):
        for a, b in zip(outs[form], outs[None]):
            assert np.array_equal(a, b), form
    outs[1] = outs["post_dual"]
    ora = _oracle(name, 8, B)
    ref = np.concatenate([ora.step_pcm(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(outs[1][0] - ref).max() < WAVE_TOL


@pytest.mark.parametrize("name,B", [("dpdfnet2", 150), ("dpdfnet2_48khz_hr", 40)])
def test_gru_tc_units_per_cta(torch_cuda, name, B):
    """k_gru_tc<UC>: 64 hidden units per CTA (throughput form, 3-deep ring of 48 KB weight slabs) and 32 (latency form:
    half slabs cut out of the same packed image, 6-deep ring, twice the CTAs).  Every output column sees the same MMA
    sequence in both, so they must agree bit for bit; against the oracle through the waveform."""
    T = 4
    hop = get_spec(name).hop
    rng = np.random.default_rng(51)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    outs = {}
    for uc in (64, 32, 0):
        eng = _engine(name, 13, B)
        eng.set_option("gru_tc", 1)
        eng.set_option("gru_uc", uc)
        outs[uc] = (eng.run_pcm_host(pcm), eng.debug_tensor("emb", B), eng.debug_tensor("m", B), eng.state_export(B - 1))
        eng.close()
    for uc in (32, 0):
        for a, b in zip(outs[uc], outs[64]):
            assert np.array_equal(a, b), uc
    ora = _oracle(name, 13, B)
    ref = np.concatenate([ora.step_pcm(pcm[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(outs[32][0] - ref).max() < WAVE_TOL


@pytest.mark.parametrize("name,B,lanes", [("dpdfnet2", 130, 1), ("dpdfnet4", 700, 1), ("dpdfnet2_48khz_hr", 67, 1), ("dpdfnet2", 600, 3)])
def test_post_kernel_persistent_with_resident_weights(torch_cuda, name, B, lanes):
    """k_dprnn_post_res: one CTA per SM keeps its branch's nine weight slabs in shared memory, draws tiles from an atomic
    counter (self-resetting: several hops and blocks reuse it) and prefetches the next tile's hcat rows into registers.
    Same arithmetic per row as k_dprnn_post_tc, so bit-identical results - with more tiles than CTAs, ragged last tiles,
    and lanes running their own counters side by side."""
    T = 5
    hop = get_spec(name).hop
    rng = np.random.default_rng(61)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    outs = {}
    for res in (0, 1):
        eng = _engine(name, 9, B)
        eng.set_option("intra_tc", 1)
        eng.set_option("overlap", 0)
        eng.set_option("lanes", lanes)
        eng.set_option("post_res", res)
        outs[res] = (eng.run_pcm_host(pcm), eng.debug_tensor("xd", B), eng.debug_tensor("xe", B), eng.state_export(B - 1))
        eng.close()
    for a, b in zip(outs[1], outs[0]):
        assert np.array_equal(a, b)
    ora = _oracle(name, 9, min(B, 8))
    ref = np.concatenate([ora.step_pcm(pcm[:8, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(outs[1][0][:8] - ref).max() < WAVE_TOL


@pytest.mark.parametrize("name,B", [("dpdfnet2", 133), ("dpdfnet2_48khz_hr", 70), ("dpdfnet4", 300)])
def test_dft_on_tensor_cores(torch_cuda, name, B):
    """k_dft_tc: the framed DFT and the inverse DFT + overlap-add as FP16-split tcgen05 GEMMs over 128-stream tiles
    (the feature / mask kernels then skip their own transforms).  Against the FFMA2 transforms (1e-5: different summation
    order only) and the oracle; ragged tiles, slot indirection, graph replay over several hops (history and overlap-add
    tail carried in state), a strided / unaligned device view, and lanes."""
    torch = torch_cuda
    T = 6
    hop = get_spec(name).hop
    rng = np.random.default_rng(71)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    slots = rng.permutation(B + 4)[:B].astype(np.int32)
    outs = {}
    for tc in (0, 1):
        eng = _engine(name, 12, B + 4)
        eng.set_option("dft_tc", tc)
        a = eng.run_pcm_host(pcm)                                   # identity slots, graph replay over T hops
        eng.reset()
        b = np.concatenate([eng.step_pcm_host(pcm[:, t * hop:(t + 1) * hop], slot_ids=slots) for t in range(T)], 1)
        st = eng.state_export(int(slots[B - 1]))
        outs[tc] = (a, b, st)
        if tc:
            # unaligned rows: a device view with an odd stride and offset goes through the scalar PCM loads
            eng.reset()
            buf = torch.zeros(B, T * hop + 3, device="cuda")
            buf[:, 1:1 + T * hop] = torch.from_numpy(pcm).cuda()
            y = eng.run_pcm(buf[:, 1:1 + T * hop], out=torch.zeros(B, T * hop + 3, device="cuda")[:, 2:2 + T * hop])
            torch.cuda.synchronize()
            assert np.array_equal(y.cpu().numpy(), a)
            eng.reset()
            eng.set_option("lanes", 2)
            assert np.array_equal(eng.run_pcm_host(pcm), a)
        eng.close()
    assert np.array_equal(outs[1][0], outs[1][1])                   # slot indirection does not change the arithmetic
    assert np.abs(outs[1][0] - outs[0][0]).max() < 1e-5
    assert np.abs(outs[1][2] - outs[0][2]).max() < 1e-5 * max(1.0, float(np.abs(outs[0][2]).max()))   # mu is in dB: |state| up to 90
    ora = _oracle(name, 12, 8)
    ref = np.concatenate([ora.step_pcm(pcm[:8, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(outs[1][0][:8] - ref).max() < WAVE_TOL


@pytest.mark.parametrize("name,B", [("dpdfnet2", 300), ("dpdfnet2_48khz_hr", 70)])
def test_schedule_options_do_not_change_the_arithmetic(torch_cuda, name, B):
    """Options that only move kernels around - the df pathway conv on a forked stream behind df_conv0 with k_df_combine on
    the tail (dfp_early), the overlapped post kernel forced on where lanes are the default (overlap = 2), lanes, the GRU
    cells' unit chunk - must reproduce the default schedule bit for bit, state included, warm-up frames included."""
    T = 5
    hop = get_spec(name).hop
    rng = np.random.default_rng(77)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    flags = np.zeros(B, np.int32)
    flags[::3] = 1                                                  # DPDF_FLAG_WARMUP rows: the coefficient ring gets zeros
    outs = []
    for opts in ({}, {"dfp_early": 1}, {"dfp_early": 1, "lanes": 2}, {"overlap": 2}, {"lanes": 1, "overlap": 0}, {"gru_uc": 64}):
        eng = _engine(name, 9, B)
        for k in ("intra_tc", "post_tc", "sep_tc", "gru_tc"):
            eng.set_option(k, 1)
        for k, v in opts.items():
            eng.set_option(k, v)
        a = np.concatenate([eng.step_pcm_host(pcm[:, t * hop:(t + 1) * hop], flags=flags if t == 0 else None) for t in range(T)], 1)
        outs.append((a, eng.state_export(B - 1), eng.state_export(0)))
        eng.close()
    for o in outs[1:]:
        for x, y in zip(o, outs[0]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("intra_tc", [0, 1])
def test_lanes_match_single_chain(torch_cuda, intra_tc):
    """A batched step split into lanes (row ranges running as forked kernel chains inside one CUDA graph) must give
    exactly what the single chain gives: identity slots, permuted slot ids, graph replay and per-hop launches."""
    name, B, T = "dpdfnet2", 400, 4
    spec = get_spec(name)
    hop = spec.hop
    rng = np.random.default_rng(21)
    pcm = (rng.standard_normal((B, T * hop)) * 0.1).astype(np.float32)
    slots = rng.permutation(B + 5)[:B].astype(np.int32)
    outs = []
    for lanes in (1, 2, 3):
        eng = _engine(name, 6, B + 5)
        eng.set_option("intra_tc", intra_tc)
        eng.set_option("lanes", lanes)
        a = eng.run_pcm_host(pcm)                                   # identity slots, graph replay over T hops
        eng.reset()
        b = np.concatenate([eng.step_pcm_host(pcm[:, t * hop:(t + 1) * hop], slot_ids=slots) for t in range(T)], 1)
        st = eng.state_export(int(slots[B - 1]))
        outs.append((a, b, st))
        assert eng.kernel_launches >= min(lanes, 2) * 20            # 400 streams fill two 256-stream lanes
    eng.set_option("free_lanes", 0)                                     # lanes forked and joined inside every hop
    eng.reset()
    assert np.array_equal(eng.run_pcm_host(pcm), outs[0][0])
    for a, b, st in outs[1:]:
        assert np.array_equal(a, outs[0][0])
        assert np.array_equal(b, outs[0][1])
        assert np.array_equal(st, outs[0][2])
    assert np.array_equal(outs[0][0], outs[0][1])                   # slot indirection does not change the arithmetic
    ora = _oracle(name, 6, 4)
    ref = np.concatenate([ora.step_pcm(pcm[:4, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(outs[2][0][:4] - ref).max() < WAVE_TOL


@pytest.mark.parametrize("dfp_ps", [0, 1])
def test_state_import_export_roundtrip(torch_cuda, dfp_ps):
    eng = _engine("dpdfnet2", 9, 3)
    eng.set_option("dfp_ps", dfp_ps)            # partial-sum form: the pending sums are rebuilt from the imported ring
    F = eng.spec.freq_bins
    rng = np.random.default_rng(0)
    X = (rng.standard_normal((7, 3, F, 2)) * 10).astype(np.float32)
    for t in range(5):                      # 5 hops: ring heads are mid-cycle for L=3 and L=5
        eng.step_spec_host(X[t])
    flat = eng.state_export(2)
    other = _engine("dpdfnet2", 9, 1)
    other.set_option("dfp_ps", dfp_ps)
    other.state_import(0, flat)
    assert np.array_equal(other.state_export(0), flat)
    a = eng.step_spec_host(X[5])
    b = other.step_spec_host(X[5][2:3])
    assert np.abs(a[2] - b[0]).max() < 1e-5
    with pytest.raises(ValueError):
        other.state_import(0, flat[:-1])


def test_reset_and_errors(torch_cuda):
    eng = _engine("dpdfnet2", 1, 4)
    F = eng.spec.freq_bins
    X = (np.random.default_rng(1).standard_normal((4, F, 2)) * 10).astype(np.float32)
    first = eng.step_spec_host(X)
    eng.step_spec_host(X)
    eng.reset([1, 3])
    again = eng.step_spec_host(X)
    assert np.array_equal(again[1], first[1]) and np.array_equal(again[3], first[3])
    assert not np.array_equal(again[0], first[0])
    with pytest.raises(ValueError):
        eng.step_spec_host(np.zeros((5, F, 2), np.float32))        # B > max_streams
    with pytest.raises(ValueError):
        eng.step_spec_host(X, slot_ids=[0, 1, 2, 9])                # slot out of range
    with pytest.raises(ValueError):
        eng.step_spec_host(np.zeros((2, F + 1, 2), np.float32))


def test_full_size_batch_invariance(torch_cuda):
    """BASELINE configs[1] size (dpdfnet4, B=1024): every stream must behave exactly as it does alone.

    Streams share nothing but weights, so (i) identical inputs in different batch rows give
    bit-identical outputs and (ii) a stream's output does not depend on the batch size."""
    torch = torch_cuda
    name, B, T = "dpdfnet4", 1024, 6
    eng = _engine(name, 0, B)
    hop = eng.spec.hop
    rng = np.random.default_rng(4)
    base = (rng.standard_normal((4, T * hop)) * 0.1).astype(np.float32)
    pcm = np.tile(base, (B // 4, 1))
    out = eng.run_pcm_host(pcm)
    assert np.isfinite(out).all()
    for r in range(4):
        assert np.array_equal(out[r::4], np.broadcast_to(out[r], (B // 4, T * hop)))
    small = _engine(name, 0, 4)
    out_small = small.run_pcm_host(base)
    assert np.abs(out_small - out[:4]).max() < 1e-5
    ora = _oracle(name, 0, 4)
    ref = np.concatenate([ora.step_pcm(base[:, t * hop:(t + 1) * hop]) for t in range(T)], 1)
    assert np.abs(out[:4] - ref).max() < WAVE_TOL
    assert eng.kernel_launches > 0
