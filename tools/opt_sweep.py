#!/usr/bin/env python
"""Kernel-time table of one hop for a list of engine option settings (GPU box).

    python tools/opt_sweep.py --model dpdfnet4 --batch 8192 16384 --opt post_pf=0 --opt post_pf=2 --opt "sep_tma=0"

Each --opt is a comma-separated list of key=value engine options applied together; prints ms/hop of a 30-hop
device-resident run (graph replay, engine-default lanes) and the per-kernel table of one un-graphed chain.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpdfnet4")
    ap.add_argument("--batch", type=int, nargs="+", default=[8192])
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--hops", type=int, default=30)
    a = ap.parse_args()
    import torch
    from dpdfnet_b200.engine import Engine
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import random_checkpoint
    spec = get_spec(a.model)
    ck = random_checkpoint(spec, 0)
    for B in a.batch:
        x = torch.from_numpy(np.clip(np.random.default_rng(1).standard_normal((B, (a.hops + 5) * spec.hop), dtype=np.float32) * 0.1, -1, 1)).cuda()
        y = torch.empty_like(x)
        for opt in (a.opt or [""]):
            eng = Engine(spec, ck, max_streams=B)
            for kv in filter(None, opt.split(",")):
                k, v = kv.split("=")
                eng.set_option(k, int(v))
            eng.run_pcm(x[:, :5 * spec.hop], out=y[:, :5 * spec.hop])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run_pcm(x[:, 5 * spec.hop:], out=y[:, 5 * spec.hop:])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.hops
            kt = eng.time_kernels(B, iters=3)
            eng.poll_error()
            top = ", ".join(f"{k} {v:.3f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1])[:7])
            print(f"{a.model} B={B} [{opt or 'default'}]: {ms:.3f} ms/hop = {B / ms / 1e3:.0f}k sf/s | {top}", flush=True)
            eng.close()


if __name__ == "__main__":
    main()
