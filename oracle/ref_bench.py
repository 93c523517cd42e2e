"""CPU baseline: the UNMODIFIED reference streaming path timed on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by ``bench.py --impl reference`` and by the ``cpu_baseline`` leg of the
GPU arm; never imported by the product package.

What runs in every worker process: the reference's public ``StreamEnhancer.process`` (package/src/dpdfnet/
stream.py:74-165 - causal framing, ``np.fft.rfft``, ``session.run``, ``irfft``, overlap-add) on the reference's
per-frame graph (onnx_model DPDFNet behind the export wrapper, export_dpdfnet_to_onnx.py:14-25), loaded from the
mounted reference tree or from ``oracle/_ref`` (verbatim copy, oracle/build_ref.py).  onnxruntime and the .onnx
files are not in the image, so the graph is executed by torch eager on ONE thread per worker
(``torch.set_num_threads(1)``, mirroring the reference's 1 intra-op / 1 inter-op thread session,
onnx_backend.py:26-28) - the fall-back SURVEY.md section 8(d) names.  One worker per host core, one stream per
worker at a time, timed like the reference's own harness (``perf_counter`` around the per-frame call,
``avg_frame_ms`` and RTF, onnx_model/infer_dpdfnet_onnx.py:99-107, 299-306).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _worker(idx, model, seed, warm, hops, barrier, q):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    try:
        import numpy as np
        import torch
        torch.set_num_threads(1)
        try:
            torch.set_num_interop_threads(1)
        except Exception:
            pass
        from dpdfnet_b200.spec import get_spec
        from dpdfnet_b200.weights import random_checkpoint
        from oracle import ref_import
        spec = get_spec(model)
        se = ref_import.reference_stream_enhancer(spec, random_checkpoint(spec, seed))
        rng = np.random.default_rng(1234 + idx)
        hop = spec.hop
        pcm = np.clip(rng.standard_normal((warm + hops + 1) * hop).astype(np.float32) * np.float32(0.1), -1, 1)
        se.process(pcm[:hop])                                    # buffered, no frame yet (stream.py:116)
        for t in range(warm):
            se.process(pcm[(t + 1) * hop:(t + 2) * hop])
        barrier.wait()
        infer = 0.0
        n_out = 0
        t_start = time.perf_counter()
        for t in range(warm, warm + hops):
            t0 = time.perf_counter()
            y = se.process(pcm[(t + 1) * hop:(t + 2) * hop])
            infer += time.perf_counter() - t0
            n_out += y.size
        wall = time.perf_counter() - t_start
        assert n_out == hops * hop, (n_out, hops * hop)
        q.put((idx, wall, infer, hops, None))
    except Exception as exc:          # surface the failure to the parent instead of hanging the barrier
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, 0.0, 0.0, 0, f"{type(exc).__name__}: {exc}"))


def run(model: str, hops: int, warm: int = 3, workers: int | None = None, seed: int = 0) -> dict:
    """Aggregate stream-frames/s of `workers` single-thread reference streams running side by side."""
    from oracle import ref_import
    if not ref_import.available():
        raise RuntimeError("reference sources not available (neither /root/reference nor oracle/_ref): run oracle/build_ref.py")
    from dpdfnet_b200.spec import get_spec
    spec = get_spec(model)
    n = workers or host_cores()
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(n)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(i, model, seed, warm, hops, barrier, q), daemon=True) for i in range(n)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    errs = [r[4] for r in res if r[4]]
    if errs:
        raise RuntimeError(f"reference worker failed: {errs[0]}")
    wall = max(r[1] for r in res)
    frames = sum(r[3] for r in res)
    avg_frame_ms = 1e3 * sum(r[2] for r in res) / frames
    hop_s = spec.hop / spec.sample_rate
    return {"value": frames / wall, "cores": n, "wall_s": wall, "frames": frames,
            "per_core": frames / wall / n, "avg_frame_ms": avg_frame_ms, "rtf": (avg_frame_ms * 1e-3) / hop_s,
            "source": ref_import.kind(),
            "what": "reference StreamEnhancer.process (stream.py:74-165) on the reference torch per-frame graph "
                    "(onnx_model DPDFNet behind DPDFNetOnnxWrapper), torch eager, 1 thread per worker process, "
                    "1 worker per host core; ORT + .onnx unavailable offline (SURVEY 8d fall-back)"}


if __name__ == "__main__":
    import json
    sys.path.insert(0, str(ROOT))
    print(json.dumps(run(sys.argv[1] if len(sys.argv) > 1 else "dpdfnet4", int(sys.argv[2]) if len(sys.argv) > 2 else 20)))
