#!/usr/bin/env python
"""Benchmark of the DPDFNet per-frame hot path on B200 (BASELINE.json contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1|cfg2|cfg3|cfg4]

A *step* is one 10 ms hop of the whole hot path (analysis STFT -> ... -> iSTFT/OLA) for one batch of synthetic
white-noise streams per GPU.  ``value`` = stream-frames/s over all ranks with inputs resident in HBM; ``e2e`` = the
same through the host-buffer C-ABI call (H2D + step + D2H per hop).

Workloads (BASELINE.json ``configs``; configs[0] is the correctness-only clip, covered by tests/):
  cfg1  configs[1]  dpdfnet4 16 kHz, 1024 streams per GPU (weak scaling)             <- default, the quoted metric
  cfg2  configs[2]  dpdfnet8 16 kHz, 4096 streams in total split over the GPUs (strong scaling sweep)
  cfg3  configs[3]  dpdfnet2_48khz_hr, 2048 streams per GPU (weak)
  cfg4  configs[4]  dpdfnet8_48khz_hr through the public streaming API (StreamGroup on the shared engine), per-tick
                    latency histogram at the largest concurrency whose p99 stays inside the 10 ms hop
The default run reports cfg1 in full and appends compact results of cfg2..cfg4 (``other_configs``) so one driver run
records every BASELINE config; ``--config cfgN`` makes any of them the headline of the line.

``--impl reference`` times the UNMODIFIED reference streaming path (its public StreamEnhancer.process on its own
per-frame torch graph, from /root/reference or the verbatim copy oracle/_ref) on the host cores: one single-thread
worker per core (oracle/ref_bench.py).  ONNX Runtime and the .onnx files are not in the image, so the graph runs under
torch eager - the fall-back SURVEY.md section 8(d) names; the numpy port of the same path is reported next to it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from dpdfnet_b200.spec import get_spec  # noqa: E402
from dpdfnet_b200.weights import pack_tensors, random_checkpoint  # noqa: E402

METRIC = "stream_frames_per_s"
UNIT = "stream-frames/s"

CONFIGS = {
    "cfg1": {"model": "dpdfnet4", "batch": 1024, "scaling": "weak", "baseline": "configs[1]"},
    "cfg2": {"model": "dpdfnet8", "batch": 4096, "scaling": "strong", "baseline": "configs[2]"},
    "cfg3": {"model": "dpdfnet2_48khz_hr", "batch": 2048, "scaling": "weak", "baseline": "configs[3]"},
    "cfg4": {"model": "dpdfnet8_48khz_hr", "batch": 2048, "scaling": "weak", "baseline": "configs[4]", "api": "StreamGroup"},
}


def bench_config(args, world: int) -> dict:
    """The ``config`` object of the JSON line - identical for the GPU arm and the reference arm of one workload."""
    c = CONFIGS[args.config]
    spec = get_spec(args.model)
    per_gpu = args.batch // world if c["scaling"] == "strong" else args.batch
    how = (f"{args.batch} streams in total split over the GPUs" if c["scaling"] == "strong" else f"batch={args.batch} streams/GPU")
    return {"workload": f"{args.model} {spec.sample_rate // 1000} kHz per-frame hot path (STFT->DPRNN->DF->iSTFT), {how}",
            "baseline_config": c["baseline"], "streams_per_gpu": per_gpu,
            "parallelism": "independent streams sharded over the GPUs, no data-path collective",
            "l2": f"no flush: per-step working set {per_gpu * 4 * spec.state_size / 1e6:.0f} MB of stream state per GPU (L2 is 126 MB); "
                  "every hop reads new PCM",
            "weights": "seeded random (no checkpoint offline), BN stats randomised; parity on shipped checkpoints not demonstrable offline",
            "input": "white noise, 0.1 RMS, clipped to +-1"}


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, source) -- the driver-measured roofline denominators."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, model: str, batch: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py traffic); None when no capture matches."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None, None
    for rec in json.loads(p.read_text()):
        if rec["kernel"] == kernel and rec["model"] == model and rec["batch"] == batch:
            return rec["dram_bytes_per_launch"], rec["source"]
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_pcm(B: int, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.clip(rng.standard_normal((B, n), dtype=np.float32) * np.float32(0.1), -1, 1)


# ------------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; give the CPU legs every host core and return the count used."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        pools = threadpoolctl.threadpool_info()
        used = max([p_.get("num_threads", 1) for p_ in pools] or [1])
        return int(used)
    except Exception:
        return 1


def cpu_port_throughput(spec, ck, B_cpu: int, hops: int, warm: int = 2):
    """stream-frames/s of the oracle port on the host cores (numpy + OpenBLAS threads)."""
    from oracle.oracle_np import OracleEngine
    ora = OracleEngine(spec, pack_tensors(spec, ck), B_cpu)
    pcm = synth_pcm(B_cpu, (hops + warm) * spec.hop, 99)
    for t in range(warm):
        ora.step_pcm(pcm[:, t * spec.hop:(t + 1) * spec.hop])
    t0 = time.perf_counter()
    for t in range(warm, warm + hops):
        ora.step_pcm(pcm[:, t * spec.hop:(t + 1) * spec.hop])
    dt = time.perf_counter() - t0
    return B_cpu * hops / dt, dt


def cpu_reference(model: str, hops: int, warm: int) -> dict:
    """The unmodified reference streaming path, one single-thread worker per host core (oracle/ref_bench.py)."""
    from oracle import ref_bench
    r = ref_bench.run(model, hops=hops, warm=warm)
    sample = (f"{r['cores']} streams (1 per core) x {hops} hops of {model} through the reference StreamEnhancer.process, "
              f"{r['wall_s']:.1f} s wall per worker")
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": sample,
            "engine": "reference-torch (torch eager, 1 thread/worker; onnxruntime + .onnx not in the image)",
            "per_core": r["per_core"], "avg_frame_ms": r["avg_frame_ms"], "rtf": r["rtf"], "source": r["source"],
            "what": r["what"]}


def cpu_baseline_block(args, spec, ck) -> dict:
    """cpu_baseline of the GPU arm: the reference when its sources travelled, the numpy port beside it."""
    out = None
    try:
        out = cpu_reference(args.model, args.cpu_hops, 3)
    except Exception as exc:  # reference sources absent: fall back to the port, say so
        out = {"kind": "port", "unit": UNIT, "reference_unavailable": f"{type(exc).__name__}: {exc}"}
    cores = use_all_host_threads()
    pv, pdt = cpu_port_throughput(spec, ck, args.cpu_batch, args.cpu_hops)
    port = {"value": pv, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{args.cpu_batch} streams x {args.cpu_hops} hops of {args.model}, batched numpy + OpenBLAS ({pdt:.1f} s)"}
    if out.get("kind") == "port":
        out.update(port)
    else:
        out["port"] = port
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    spec = get_spec(args.model)
    ck = random_checkpoint(spec, 0)
    try:
        cb = cpu_reference(args.model, args.steps, args.warmup)
    except Exception as exc:
        cores = use_all_host_threads()
        pv, pdt = cpu_port_throughput(spec, ck, args.cpu_batch, args.steps, warm=args.warmup)
        cb = {"value": pv, "unit": UNIT, "cores": cores, "kind": "port", "reference_unavailable": f"{type(exc).__name__}: {exc}",
              "sample": f"{args.cpu_batch} streams x {args.steps} hops of {args.model}, batched numpy + OpenBLAS ({pdt:.1f} s)"}
    else:
        cores = use_all_host_threads()
        pv, pdt = cpu_port_throughput(spec, ck, args.cpu_batch, min(args.steps, 20))
        cb["port"] = {"value": pv, "unit": UNIT, "cores": cores, "kind": "port",
                      "sample": f"{args.cpu_batch} streams x {min(args.steps, 20)} hops, batched numpy + OpenBLAS ({pdt:.1f} s)"}
    val = cb["value"]
    fps = spec.sample_rate / spec.hop
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * cb["cores"] / val if cb["kind"] == "reference" else 1e3 * args.cpu_batch / val,
            "higher_is_better": True, "scaling": CONFIGS[args.config]["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world), "cpu_baseline": cb,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "realtime_streams": val / fps, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing of one rank (NCCL only for barriers and the max-over-ranks timer)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            self.dist = dist

    def sync(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, v: float) -> float:
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_run(D: Dist, eng, pcm, out, W: int, K: int, hop: int) -> float:
    """W warm-up hops, then K hops between barriers + synchronize; CUDA events on the launching stream; max over ranks (ms)."""
    torch = D.torch
    stream = torch.cuda.current_stream()
    eng.run_pcm(pcm[:, :W * hop], out=out[:, :W * hop])
    D.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.sync()
    ev0.record(stream)
    eng.run_pcm(pcm[:, W * hop:(W + K) * hop], out=out[:, W * hop:(W + K) * hop])
    ev1.record(stream)
    D.sync()
    return D.max(ev0.elapsed_time(ev1))


def lockstep_latency(D: Dist, eng, pcm, out, hops: int, hop: int) -> dict:
    """Device time of every single hop (one graph launch per hop from the host, CUDA events around each)."""
    torch = D.torch
    stream = torch.cuda.current_stream()
    nin = pcm.shape[1] // hop
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(hops + 1)]
    D.sync()
    evs[0].record(stream)
    for t in range(hops):
        i = t % nin
        eng.step_pcm(pcm[:, i * hop:(i + 1) * hop], out=out[:, :hop])
        evs[t + 1].record(stream)
    torch.cuda.synchronize()
    ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(hops)])
    p50, p99, mx = float(np.percentile(ms, 50)), float(np.percentile(ms, 99)), float(ms.max())
    return {"p50_ms": D.max(p50), "p99_ms": D.max(p99), "max_ms": D.max(mx), "hops": hops}


def sustained_ladder(D: Dist, spec, ck, batches, hops: int):
    """Largest batch per GPU whose mean AND p99 lock-step hop latency stay inside the hop period, all ranks loaded."""
    from dpdfnet_b200.engine import Engine
    torch = D.torch
    hop = spec.hop
    fps = spec.sample_rate / hop
    ladder = []
    for Bl in batches:
        engl = Engine(spec, ck, max_streams=Bl, device=D.local)
        xl = torch.from_numpy(synth_pcm(Bl, 30 * hop, 777 + D.rank)).cuda()
        yl = torch.empty_like(xl)
        ms = timed_run(D, engl, xl, yl, 5, 25, hop) / 25.0
        lat = lockstep_latency(D, engl, xl, yl, hops, hop)
        rec = {"streams_per_gpu": Bl, "ms_per_hop": ms, "stream_frames_per_s": D.world * Bl / (ms * 1e-3), "hop_latency_ms": lat}
        if D.rank == 0:
            ktl = engl.time_kernels(Bl, iters=2)
            rec["kernel_ms"] = {k: round(v, 4) for k, v in sorted(ktl.items(), key=lambda kv: -kv[1])}
            rec["tensor_tflops"] = tensor_rates(spec, ktl, Bl)
        ladder.append(rec)
        engl.close()
        del engl, xl, yl
    ok = [r for r in ladder if max(r["ms_per_hop"], r["hop_latency_ms"]["p99_ms"]) < 1e3 / fps]
    best = max(ok, key=lambda r: r["streams_per_gpu"]) if ok else None
    return {"ladder": ladder, "hop_budget_ms": 1e3 / fps, "latency_hops_per_batch": hops,
            "realtime_streams_per_gpu": best["streams_per_gpu"] if best else None,
            "realtime_streams_total": D.world * best["streams_per_gpu"] if best else None,
            "ms_per_hop_at_that_batch": best["ms_per_hop"] if best else None,
            "p99_hop_latency_ms_at_that_batch": best["hop_latency_ms"]["p99_ms"] if best else None,
            "criterion": "largest ladder batch whose mean AND p99 lock-step hop latency (max over ranks) are below the hop period"}


def kernel_models(spec):
    """Algorithmic bytes and MACs per stream per LAUNCH of the kernels that can dominate a hop."""
    Fe3, Fd = spec.fe[3], 48
    fe = spec.fe
    # separable convs: activations in (+ pathway source) and out, per launch group as enqueue_step issues them
    sep_rows_in = [96 * 6 / 64 + fe[0], 96 + fe[1], fe[2], 2 * fe[3], 2 * fe[2], 2 * fe[1]]      # [F][64] rows read (df ring counted as rows)
    sep_rows_out = [2 * 96 + fe[1], 48 + fe[2], fe[3], fe[2], fe[1], fe[0]]
    sep_bytes = 4 * 64 * (sum(sep_rows_in) + sum(sep_rows_out)) / 6.0
    sep_macs = 64 * 64 * sum([96 + fe[1], 48 + fe[2], fe[3], fe[2], fe[1], fe[0]]) / 6.0
    return {
        "dprnn_intra": {"bytes": 4 * (Fe3 + Fd) * (64 + 128), "macs": (Fe3 + Fd) * 2 * 2 * 192 * 64},
        "dprnn_post": {"bytes": 4 * (Fe3 + Fd) * (128 + 64 + 64 + 2 * 64), "macs": (Fe3 + Fd) * (128 * 64 + 6 * 64 * 64 + 64 * 64)},
        "sepconv": {"bytes": sep_bytes, "macs": sep_macs},
    }


def tensor_rates(spec, kt, B):
    """FP32-equivalent TFLOP/s of the tensor-core kernels at this batch (x3 = issued FP16 tensor FLOP/s)."""
    km = kernel_models(spec)
    out = {}
    for k, n in (("dprnn_intra", spec.n_blocks), ("dprnn_post", spec.n_blocks), ("sepconv", 6)):
        if k in kt and n:
            out[k] = 2.0 * km[k]["macs"] * B / (kt[k] / n * 1e-3) / 1e12
    return out


def roofline_blocks(spec, args, kt, B, ms_per_step):
    peak, peak_tf, peak_src = measured_peaks()
    km = kernel_models(spec)
    step_ms_sum = sum(kt.values())
    dom = max(kt, key=kt.get)
    n_dom = {"dprnn_intra": spec.n_blocks, "dprnn_post": spec.n_blocks, "sepconv": 6, "gl": 6, "gru": 3}.get(dom, 1)
    per_bytes = km.get(dom, {}).get("bytes", spec.algorithmic_bytes_per_frame)
    per_macs = km.get(dom, {}).get("macs")
    dom_ms = kt[dom] / n_dom
    achieved = per_bytes * B / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom, args.model, B)
    dom_tf = 2.0 * per_macs * B / (dom_ms * 1e-3) / 1e12 if per_macs else None
    step_bytes = spec.algorithmic_bytes_per_frame
    step_gbs = step_bytes * B / (ms_per_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "share_of_step": kt[dom] / step_ms_sum, "launch_ms": dom_ms, "launches_per_step": n_dom,
            "algorithmic_bytes_per_stream_launch": per_bytes,
            "note": "at this batch the hop is bound by the latency of its dependent chain (sequential GRU steps), not by HBM: "
                    "activations stay in L2 (traffic << algorithmic bytes); roofline_tensor and sustained give the throughput regime"}
    roof_t = None if dom_tf is None else {
        "bound": "tensor", "kernel": dom, "achieved": dom_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": dom_tf / peak_tf,
        "issued_frac": 3.0 * dom_tf / peak_tf, "algorithmic_macs_per_stream_launch": per_macs,
        "note": "FP32-equivalent algorithmic FLOPs; every MAC is issued as 3 FP16 tensor MACs (hi*hi + lo*hi + hi*lo); peak = measured dense bf16 burst. "
                "issued_frac counts those 3 passes only: below 3 073 streams the df-branch sweep spreads every stream over D = 2 or 4 rows of "
                "the M = 128 tile (same MMAs, identical rows), tensor work that buys latency and is not counted here; up to 1 792 streams it runs "
                "in fragment form instead - 32 streams on 64 of the 128 rows (hi | lo operand halves), TWO passes per product"}
    roof_s = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
              "algorithmic_bytes_per_stream_frame": step_bytes}
    return roof, roof_t, roof_s


def run_engine_config(D: Dist, args, model: str, B: int, K: int, W: int, full: bool):
    """Device-resident throughput (+ e2e, latency histogram, kernel table when `full`) of `model` at B streams per GPU."""
    from dpdfnet_b200.engine import Engine
    torch = D.torch
    spec = get_spec(model)
    ck = random_checkpoint(spec, 0)
    hop = spec.hop
    eng = Engine(spec, ck, max_streams=B, device=D.local)
    if full:
        if args.no_graph:
            eng.set_option("graph", 0)
        if args.lanes >= 0:
            eng.set_option("lanes", args.lanes)
        for kv in args.opt:
            k, v = kv.split("=")
            eng.set_option(k, int(v))
    if D.world > 1:
        D.sync()
    pcm = torch.from_numpy(synth_pcm(B, (W + K) * hop, 1234 + D.rank)).cuda()
    out = torch.empty_like(pcm)
    sampler = ClockSampler(D.local)
    if D.rank == 0 and full:
        sampler.start()
    ms = timed_run(D, eng, pcm, out, W, K, hop)
    clocks = sampler.stop() if (D.rank == 0 and full) else None
    res = {"model": model, "streams_per_gpu": B, "ms_per_step": ms / K, "value": D.world * B * K / (ms * 1e-3),
           "launches_per_step": eng.kernel_launches, "clocks": clocks, "spec": spec, "ck": ck}
    if args.profile_only:
        eng.close()
        return res
    # ---- end to end: host buffers through the C ABI, copies inside the timed region
    Ke = min(K, args.e2e_steps)
    pin_in = torch.from_numpy(np.ascontiguousarray(
        synth_pcm(B, Ke * hop, 4321 + D.rank).reshape(B, Ke, hop).transpose(1, 0, 2))).pin_memory()    # [Ke][B][hop]: one pinned block per hop
    pin_out = torch.empty(B, hop).pin_memory()
    host_in, host_out = pin_in.numpy(), pin_out.numpy()
    eng.reset()
    for t_ in range(min(3, Ke)):
        eng.step_pcm_host(host_in[t_], out=host_out)
    D.sync()
    t0 = time.perf_counter()
    sink = 0.0
    for t_ in range(Ke):
        y = eng.step_pcm_host(host_in[t_], out=host_out)
        sink += float(y[0, 0])
    torch.cuda.synchronize()
    sync_s = D.max(time.perf_counter() - t0)
    # the same Ke hops through the pipelined host entry: every hop's H2D copy, kernels and D2H read are still inside the
    # timed region, but hop t+1 is submitted before hop t is collected (two pinned result buffers), as a server with a
    # steady packet stream does; each result is read on the host
    pin_out2 = torch.empty(2, B, hop).pin_memory()
    host_out2 = pin_out2.numpy()
    eng.reset()
    for t_ in range(min(3, Ke)):
        eng.wait(eng.submit_pcm_host(host_in[t_], host_out2[t_ & 1]))
    D.sync()
    t0 = time.perf_counter()
    prev = None
    for t_ in range(Ke):
        tk = eng.submit_pcm_host(host_in[t_], host_out2[t_ & 1])
        if prev is not None:
            eng.wait(prev)
            sink += float(host_out2[(t_ - 1) & 1][0, 0])
        prev = tk
    eng.wait(prev)
    sink += float(host_out2[(Ke - 1) & 1][0, 0])
    e2e_s = D.max(time.perf_counter() - t0)
    res["e2e"] = {"value": D.world * B * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * hop * 4, "d2h_bytes_per_step": B * hop * 4,
                  "steps": Ke, "api": "dpdf_submit_pcm_host / dpdf_wait (pinned host buffers; H2D + hop + D2H of every hop, two hops in flight)",
                  "synchronous_value": D.world * B * Ke / sync_s,
                  "synchronous_api": "dpdf_step_pcm_host (H2D + hop + D2H + wait, one hop at a time)"}
    # ---- per-hop latency distribution at this batch
    eng.reset()
    eng.run_pcm(pcm[:, :W * hop], out=out[:, :W * hop])
    lat = lockstep_latency(D, eng, pcm, out, args.latency_hops if full else min(args.latency_hops, 200), hop)
    lat["batch"] = B
    lat["note"] = "device time per hop, one graph launch per hop from the host; budget is the 10 ms hop"
    res["hop_latency"] = lat
    if D.rank == 0:
        eng.reset()
        kt = eng.time_kernels(B, iters=5 if full else 2)
        res["kernel_ms"] = {k: round(v, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1])}
        res["kt"] = kt
    eng.close()
    return res


def run_stream_api(D: Dist, args, model: str, batches, ticks: int):
    """BASELINE configs[4]: the public streaming API (StreamGroup = B StreamEnhancer-equivalent streams on the shared engine),
    one tick = every stream hands in one hop of HOST audio and gets one hop back.  Wall-clock latency per tick
    (H2D + hop + D2H + Python), histogram at the largest concurrency whose p99 stays inside the hop period."""
    from dpdfnet_b200.onnx_backend import EnginePool
    from dpdfnet_b200.stream import StreamGroup
    torch = D.torch
    spec = get_spec(model)
    hop = spec.hop
    fps = spec.sample_rate / hop
    os.environ["DPDFNET_B200_RANDOM_WEIGHTS"] = "1"
    os.environ.pop("DPDFNET_MODEL_DIR", None)
    rows = []
    for B in batches:
        EnginePool.shutdown()
        g = StreamGroup(model=model, streams=B, device=D.local)
        x = synth_pcm(B, 8 * hop, 555 + D.rank)
        for t in range(10):                                  # first window + graph capture + warm-up
            g.process(x[:, (t % 8) * hop:(t % 8 + 1) * hop], spec.sample_rate)
        D.sync()
        lat = np.empty(ticks)
        for t in range(ticks):
            t0 = time.perf_counter()
            y = g.process(x[:, (t % 8) * hop:(t % 8 + 1) * hop], spec.sample_rate)
            lat[t] = (time.perf_counter() - t0) * 1e3
        assert y.shape == (B, hop)
        g.close()
        hist, edges = np.histogram(lat, bins=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 20, 1e9])
        rows.append({"streams_per_gpu": B, "p50_ms": D.max(float(np.percentile(lat, 50))), "p99_ms": D.max(float(np.percentile(lat, 99))),
                     "max_ms": D.max(float(lat.max())), "mean_ms": D.max(float(lat.mean())), "ticks": ticks,
                     "histogram_ms_edges": [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 20, "inf"], "histogram_counts_rank0": hist.tolist(),
                     "stream_frames_per_s": D.world * B / (D.max(float(lat.mean())) * 1e-3)})
    # the same tick through N separate StreamEnhancer objects (reference API, one object per stream) and process_many:
    # one batched engine call, but per-object Python bookkeeping on the host
    many = []
    from dpdfnet_b200.stream import StreamEnhancer, process_many
    from dpdfnet_b200.onnx_backend import reserve
    for N in args.many_ladder:
        EnginePool.shutdown()
        reserve(model, N, device=D.local)
        os.environ["DPDFNET_B200_DEVICE"] = str(D.local)
        es = [StreamEnhancer(model=model) for _ in range(N)]
        x = synth_pcm(N, 8 * hop, 99 + D.rank)
        for t in range(6):
            process_many(es, list(x[:, (t % 8) * hop:(t % 8 + 1) * hop]), spec.sample_rate)
        lat = np.empty(min(ticks, 100))
        for t in range(lat.size):
            chunks = list(x[:, (t % 8) * hop:(t % 8 + 1) * hop])
            t0 = time.perf_counter()
            ys = process_many(es, chunks, spec.sample_rate)
            lat[t] = (time.perf_counter() - t0) * 1e3
        assert len(ys) == N and ys[0].shape == (hop,)
        for e_ in es:
            e_.close()
        many.append({"enhancers": N, "p50_ms": D.max(float(np.percentile(lat, 50))), "p99_ms": D.max(float(np.percentile(lat, 99))),
                     "ticks": int(lat.size)})
    EnginePool.shutdown()
    ok = [r for r in rows if r["p99_ms"] < 1e3 / fps]
    best = max(ok, key=lambda r: r["streams_per_gpu"]) if ok else None
    return {"model": model, "api": "dpdfnet_b200.stream.StreamGroup.process (host numpy in/out, pinned staging, shared EnginePool engine)",
            "hop_budget_ms": 1e3 / fps, "ladder": rows,
            "max_concurrent_streams_per_gpu": best["streams_per_gpu"] if best else None,
            "max_concurrent_streams_total": D.world * best["streams_per_gpu"] if best else None,
            "latency_at_max": best,
            "process_many": {"api": "dpdfnet_b200.stream.process_many over separate StreamEnhancer objects sharing one pooled engine", "ladder": many}}


def run_ours(args):
    D = Dist()
    torch = D.torch
    cfg = CONFIGS[args.config]
    K, W = args.steps, args.warmup
    B = args.batch // D.world if cfg["scaling"] == "strong" else args.batch
    if B <= 0:
        raise SystemExit("batch smaller than the number of GPUs")
    spec = get_spec(args.model)
    fps = spec.sample_rate / spec.hop

    main = run_engine_config(D, args, args.model, B, K, W, full=True)
    if args.profile_only:
        if D.rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": main["ms_per_step"], "value_under_profiler": main["value"]}))
        D.close()
        return 0
    ck = main.pop("ck")
    main.pop("spec")

    stream_api = None
    if cfg.get("api") == "StreamGroup" or (args.config == "cfg1" and not args.no_extras):
        m4 = CONFIGS["cfg4"]["model"]
        stream_api = run_stream_api(D, args, m4 if args.config == "cfg1" else args.model, args.stream_ladder, args.stream_ticks)

    sustained = None
    if not args.no_ladder and args.config == "cfg1":
        sustained = sustained_ladder(D, spec, ck, args.ladder, args.latency_hops)

    others = []
    if args.config == "cfg1" and not args.no_extras:
        for name in ("cfg2", "cfg3"):
            c = CONFIGS[name]
            Bo = c["batch"] // D.world if c["scaling"] == "strong" else c["batch"]
            r = run_engine_config(D, args, c["model"], Bo, min(K, 40), W, full=False)
            sp = r.pop("spec"); r.pop("ck"); r.pop("clocks")
            kt = r.pop("kt", None)
            rec = {"config": name, "baseline_config": c["baseline"], "scaling": c["scaling"], "model": c["model"], "streams_per_gpu": Bo,
                   "streams_total": Bo * D.world, "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": UNIT,
                   "realtime_streams": r["value"] / (sp.sample_rate / sp.hop), "e2e_value": r["e2e"]["value"],
                   "hop_latency": r["hop_latency"], "kernel_ms": r.get("kernel_ms")}
            if kt is not None:
                rf, rt, rs = roofline_blocks(sp, argparse.Namespace(model=c["model"]), kt, Bo, r["ms_per_step"])
                rec["roofline"] = {k: rf[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "share_of_step", "launch_ms")}
                rec["roofline_tensor"] = None if rt is None else {k: rt[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "issued_frac")}
            others.append(rec)

    if D.rank != 0:
        D.close()
        return 0

    kt = main.pop("kt")
    roof, roof_t, roof_s = roofline_blocks(spec, args, kt, B, main["ms_per_step"])
    flops = 2.0 * spec.macs_per_frame
    cpu = cpu_baseline_block(args, spec, ck)
    e2e = main["e2e"]
    if sustained:
        e2e["realtime_streams_total"] = sustained["realtime_streams_total"]
        e2e["p99_hop_latency_ms_at_realtime_streams"] = sustained["p99_hop_latency_ms_at_that_batch"]
    config = bench_config(args, D.world)
    line = {
        # bulky detail first, the contract keys and the one-glance summary last (drivers keep the tail of the line)
        "sustained": sustained,
        "stream_api": stream_api,
        "other_configs": others or None,
        "kernel_ms": main["kernel_ms"],
        "hop_latency": main["hop_latency"],
        "roofline_tensor": roof_t,
        "roofline_step": roof_s,
        "fp32_tflops": flops * main["value"] / 1e12,
        "engine": "one CUDA graph per hop; DPRNN, separable convs and GRU(256) on tcgen05 (FP16 hi/lo split, FP32 accumulate in TMEM)",
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": D.world, "steps": K, "warmup": W,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "gpu_launches": main["launches_per_step"] * K,
        "roofline": roof, "cpu_baseline": cpu, "clocks": main["clocks"], "e2e": e2e,
        "realtime_streams": main["value"] / fps,
        "hop_latency_ms": main["ms_per_step"],
        "p50_hop_latency_ms": main["hop_latency"]["p50_ms"],
        "p99_hop_latency_ms": main["hop_latency"]["p99_ms"],
        "realtime_streams_total": sustained["realtime_streams_total"] if sustained else None,
        "realtime_streams_per_gpu": sustained["realtime_streams_per_gpu"] if sustained else None,
        "p99_hop_latency_ms_at_realtime_streams": sustained["p99_hop_latency_ms_at_that_batch"] if sustained else None,
        "stream_api_max_concurrent_streams_total": stream_api["max_concurrent_streams_total"] if stream_api else None,
    }
    print(json.dumps(line))
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg1", choices=sorted(CONFIGS))
    ap.add_argument("--model", default=None, help="override the config's model")
    ap.add_argument("--batch", type=int, default=None, help="override the config's batch (per GPU if weak, total if strong)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-batch", type=int, default=128)
    ap.add_argument("--cpu-hops", type=int, default=20)
    ap.add_argument("--latency-hops", type=int, default=500, help="lock-step hops per latency histogram")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-ladder", action="store_true", help="skip the sustained-streams batch ladder")
    ap.add_argument("--no-extras", action="store_true", help="cfg1 only: skip the compact cfg2..cfg4 results")
    ap.add_argument("--ladder", type=int, nargs="*", default=[8192, 16384, 20480, 22528, 23552, 24576, 25600, 26624])
    ap.add_argument("--stream-ladder", type=int, nargs="*", default=[1024, 2048, 3072, 4096, 4608, 5120, 5632])
    ap.add_argument("--stream-ticks", type=int, default=300)
    ap.add_argument("--many-ladder", type=int, nargs="*", default=[256, 1024], help="StreamEnhancer objects per process_many tick")
    ap.add_argument("--lanes", type=int, default=-1, help="kernel-chain lanes per step (-1: engine default)")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--profile-only", action="store_true", help="device steps only (for ncu runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    c = CONFIGS[args.config]
    args.model = args.model or c["model"]
    args.batch = args.batch or c["batch"]
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
