"""Stage-by-stage comparison of the CUDA engine against the CPU oracle (prints, never asserts).

    python tools/gpu_debug.py [model ...]

Used on the GPU box to localise a numerical discrepancy to one kernel in a single run.
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from dpdfnet_b200.engine import Engine  # noqa: E402
from dpdfnet_b200.spec import get_spec  # noqa: E402
from dpdfnet_b200.weights import pack_tensors, random_checkpoint  # noqa: E402
from oracle.oracle_np import OracleEngine  # noqa: E402

STAGES = ["e0", "e1", "e2", "e3", "c0", "hcat_e", "hcat_d", "xe", "xd", "cemb", "g0", "henc", "emb", "herb2", "ed",
          "d3", "d2", "d1", "m", "hdf2", "cc", "co"]


def oracle_stage(o: OracleEngine, name: str, N: int):
    d = o.dbg
    if name == "xe":
        return d[f"xe{N - 1}"] if N else None
    if name == "xd":
        return d[f"xd{N - 1}"] if N else d["c1"]
    if name == "hcat_e":
        return d[f"erb_hcat{N - 1}"] if N else None
    if name == "hcat_d":
        return d[f"df_hcat{N - 1}"] if N else None
    return d.get(name)


def compare(model: str, B: int = 3, frames: int = 4, seed: int = 0):
    spec = get_spec(model)
    ck = random_checkpoint(spec, seed)
    eng = Engine(spec, ck, max_streams=B + 2)
    eng.set_option("graph", 0)
    for kv in os.environ.get("DPDF_OPTIONS", "").split(","):       # e.g. DPDF_OPTIONS=intra_tc=1,post_tc=0
        if "=" in kv:
            eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    ora = OracleEngine(spec, pack_tensors(spec, ck), B + 2)
    rng = np.random.default_rng(3)
    slots = np.arange(B, dtype=np.int32)[::-1].copy() + 1     # exercise the slot indirection
    print(f"== {model}: B={B} frames={frames} launches/hop=?")
    for t in range(frames):
        X = (rng.standard_normal((B, spec.freq_bins, 2)) * 20).astype(np.float32)
        y = eng.step_spec_host(X, slot_ids=slots)
        yo = ora.step_spec(X, slots=slots.astype(np.int64))
        row = [f"t={t} out {np.abs(y - yo).max():.2e}/{np.abs(yo).max():.1f}"]
        for s in STAGES:
            ref = oracle_stage(ora, s, spec.n_blocks)
            if ref is None:
                continue
            try:
                got = eng.debug_tensor(s, B)
            except Exception as ex:  # noqa: BLE001
                row.append(f"{s}:ERR({ex})")
                continue
            ref = np.asarray(ref, np.float32).reshape(B, -1)
            if got.shape != ref.shape:
                row.append(f"{s}:shape{got.shape}!={ref.shape}")
                continue
            bad = "" if np.isfinite(got).all() else "NAN!"
            row.append(f"{s}:{np.abs(got - ref).max():.1e}{bad}")
        print("  " + " ".join(row), flush=True)
    for b in range(B):
        ds = np.abs(eng.state_export(int(slots[b])) - ora.export_state(int(slots[b])))
        off = 0
        worst = []
        for name, shape in spec.state_segments():
            n = int(np.prod(shape))
            worst.append((ds[off:off + n].max(), name))
            off += n
        w = max(worst)
        print(f"  state slot {slots[b]}: max diff {ds.max():.2e} (worst segment {w[1]})")
    print(f"  kernel launches per hop: {eng.kernel_launches}")
    # PCM path
    eng.reset()
    ora.reset()
    for t in range(frames + 4):
        pcm = (rng.standard_normal((B, spec.hop)) * 0.1).astype(np.float32)
        y = eng.step_pcm_host(pcm)
        yo = ora.step_pcm(pcm)
        print(f"  pcm t={t}: out diff {np.abs(y - yo).max():.2e} scale {np.abs(yo).max():.3f}")
    eng.close()


if __name__ == "__main__":
    models = sys.argv[1:] or ["dpdfnet2", "dpdfnet2_48khz_hr", "baseline"]
    for m in models:
        t0 = time.time()
        try:
            compare(m)
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"!! {m} failed: {ex}")
        print(f"   ({time.time() - t0:.1f}s)")
