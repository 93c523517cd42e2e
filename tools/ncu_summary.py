#!/usr/bin/env python
"""Condense ncu output into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv          > profiles/X_launches.txt
    python tools/ncu_summary.py full     gpurun_out/X_kernel.ncu-rep [...]  > profiles/X_ncu_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    tot, cnt, shape = collections.defaultdict(float), collections.Counter(), {}
    for r in rows[1:]:
        n = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        tot[n] += v
        cnt[n] += 1
        shape[n] = (r[gi], r[bi])
    s = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {s:.1f} us (ncu-serialised, cold cache: compare SHARES)")
    print(f"{'kernel':28s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>8s}  grid / block")
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{n:28s} {cnt[n]:8d} {v:10.1f} {100 * v / s:6.1f}% {v / cnt[n]:8.1f}  {shape[n][0]} / {shape[n][1]}")


def full(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"# {path}: {len(rows) - 2} profiled launch(es); first launch shown, durations of all launches listed")
        ti = hdr.index("gpu__time_duration.sum")
        print("durations:", ", ".join(f"{r[ti]} {units[ti]}" for r in rows[2:]))
        r = rows[2]
        print("kernel:", r[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            if any(h == k or (k.endswith("limit") and h.startswith(k)) for k in KEEP):
                print(f"  {h} [{units[i]}] = {r[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp-state samples:", ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:8]))
        print()


def traffic(args):
    """traffic <model> <batch> <rep...>: merge DRAM bytes per launch of every profiled kernel into profiles/ncu_traffic.json
    (read by bench.py for roofline.traffic).  Kernel names are mapped to the bench's kernel_ms keys."""
    import json
    import os
    model, batch, reps = args[0], int(args[1]), args[2:]
    names = {"k_dprnn_intra_tc": "dprnn_intra", "k_dprnn_intra": "dprnn_intra", "k_dprnn_post_tc": "dprnn_post", "k_dprnn_post": "dprnn_post",
             "k_sepconv_tc": "sepconv", "k_sepconv_tma": "sepconv", "k_sepconv": "sepconv", "k_gru_tc": "gru", "k_gl": "gl", "k_analysis": "analysis",
             "k_synthesis": "synthesis", "k_df_pathway": "df_pathway"}
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
    recs = json.load(open(out_path)) if os.path.exists(out_path) else []
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reps:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ri, wi, ki = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        by = collections.defaultdict(list)
        for r in rows[2:]:
            kn = r[ki].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
            by[kn].append(float(r[ri].replace(",", "")) * unit[units[ri]] + float(r[wi].replace(",", "")) * unit[units[wi]])
        for kn, vals in by.items():
            key = names.get(kn, kn)
            recs = [x for x in recs if not (x["kernel"] == key and x["model"] == model and x["batch"] == batch)]
            recs.append({"kernel": key, "model": model, "batch": batch, "dram_bytes_per_launch": int(sum(vals) / len(vals)),
                         "source": f"{os.path.basename(path)} ({kn}, {len(vals)} launch(es), ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"})
    json.dump(recs, open(out_path, "w"), indent=1)
    print(f"{out_path}: {len(recs)} records")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        full(sys.argv[2:])
