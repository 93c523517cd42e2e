#!/usr/bin/env python
"""In-graph cost of every kernel of a hop: hop time with only the first k kernels enqueued, k = 1 .. all (GPU box).

    python tools/chain_profile.py --model dpdfnet4 --batch 1024 [--opt key=value,...]

`dpdf_time_kernels` times each kernel alone between events; this shows what a kernel adds to the REAL chain - graph
replay, overlapped post kernel, forked decoder tails, programmatic dependent launches - by truncating the hop after k
launches (engine option `stop_after`; outputs are garbage, timing is not: no kernel's duration depends on values).
"""
import argparse
import ctypes
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpdfnet4")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--opt", default="")
    ap.add_argument("--hops", type=int, default=40)
    a = ap.parse_args()
    import torch
    from dpdfnet_b200.engine import Engine
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import random_checkpoint
    spec = get_spec(a.model)
    B = a.batch
    eng = Engine(spec, random_checkpoint(spec, 0), max_streams=B)
    for kv in filter(None, a.opt.split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    x = torch.from_numpy(np.clip(np.random.default_rng(1).standard_normal((B, a.hops * spec.hop), dtype=np.float32) * 0.1, -1, 1)).cuda()
    y = torch.empty_like(x)
    # ordered kernel names of one chain
    ms = (ctypes.c_float * 256)()
    names = (ctypes.c_char_p * 256)()
    n = ctypes.c_int32()
    eng._check(eng._lib.dpdf_time_kernels(eng._handle, B, 2, ms, names, 256, ctypes.byref(n)))
    order = [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def hop_ms():
        eng.run_pcm(x[:, :5 * spec.hop], out=y[:, :5 * spec.hop])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_pcm(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.hops

    full = hop_ms()
    print(f"{a.model} B={B} [{a.opt or 'default'}]: {full:.4f} ms/hop, {len(order)} kernels")
    print(f"{'k':>3s} {'kernel':14s} {'alone us':>9s} {'cumulative us':>14s} {'adds us':>9s}")
    prev = 0.0
    for k in range(1, len(order) + 1):
        eng.set_option("stop_after", k)
        t = hop_ms() * 1e3
        print(f"{k:3d} {order[k - 1][0]:14s} {order[k - 1][1] * 1e3:9.1f} {t:14.1f} {t - prev:9.1f}", flush=True)
        prev = t
    eng.close()


if __name__ == "__main__":
    main()
