#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the pipeline kernels into profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> <n_streams> <frames_per_launch_per_stream> <out.md> [traffic.json]
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import json
import subprocess
import sys

rep, n_streams, nf, out_md = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
traffic_json = sys.argv[5] if len(sys.argv) > 5 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    try:
        return float(r[idx[key]].replace(",", ""))
    except Exception:
        return float("nan")


def scale(r, key):  # to base units
    v, u = num(r, key), units[idx[key]]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u, 1)
    return v * mult


cols = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
traffic = {}
lines = ["| kernel | time (us) | grid x block | regs | warp-inst | inst / frame | issue active % | warps active % | "
         "DRAM read (MB) | DRAM write (MB) | top stalls (pc samples) |", "|---|---|---|---|---|---|---|---|---|---|---|"]
frames = n_streams * nf
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    t = scale(r, "gpu__time_duration.sum")
    rd, wr = scale(r, "dram__bytes_read.sum"), scale(r, "dram__bytes_write.sum")
    stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): num(r, h) for h in hdr
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(v for v in stalls.values() if v == v) or 1
    top = ", ".join(f"{k} {v / tot * 100:.0f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
    inst = num(r, "smsp__inst_executed.sum")
    lines.append(f"| {name} | {t * 1e6:.1f} | {int(num(r, 'launch__grid_size'))} x {int(num(r, 'launch__block_size'))} | "
                 f"{int(num(r, 'launch__registers_per_thread'))} | {inst / 1e6:.1f} M | {inst / frames:.0f} | "
                 f"{num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {top} |")
    traffic[name] = {"dram_bytes_per_launch": rd + wr, "streams": n_streams, "frames_per_launch": frames,
                     "gpu_time_us_isolated": t * 1e6,
                     "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "smem_wavefronts_pct_of_peak": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                     "fma_pipe_pct": num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                     "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
open(out_md, "a").write("\n".join(lines) + "\n")
if traffic_json:
    json.dump(traffic, open(traffic_json, "w"), indent=1)
print("\n".join(lines))
