#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/prof_*.ncu-rep) into profiles/ncu_summary.md.

For every captured launch: duration, DRAM traffic, achieved DRAM bandwidth as a fraction of the MEASURED copy bandwidth
(MEASURED_PEAKS.json, fallback 6576 GB/s), tensor-pipe activity, occupancy, registers, L2 hit rate and the top warp-stall
reasons.  Run here (no GPU needed): python benchmarks/ncu_summary.py
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, bf16 = 6576.4, 1689.8
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            flat = json.dumps(j)
            for k, v in (j.items() if isinstance(j, dict) else []):
                if isinstance(v, (int, float)) and "hbm" in k.lower():
                    hbm = float(v)
                if isinstance(v, (int, float)) and "bf16" in k.lower():
                    bf16 = float(v)
            del flat
        except Exception:
            pass
    return hbm, bf16


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return num(v) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    return num(v) * m.get(unit, 1)


def main():
    hbm, bf16 = peaks()
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep")))
    out = ["# ncu --set full captures (B200, clocks untouched)", "",
           f"Roofline denominators: measured copy bandwidth {hbm:.0f} GB/s, measured cuBLAS bf16 {bf16:.0f} TFLOP/s (MEASURED_PEAKS.json).", "",
           "| kernel | grid x block | regs | dur (us) | DRAM R+W (MB) | DRAM GB/s | % of measured BW | tensor pipe active % | warps active % | L2 hit % | top stalls |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled")]
        stall_cols = [h for h in stall_cols if h.endswith("_per_issue_active.ratio") or h.endswith(".ratio")]

        def g(r, name):
            return r[ix[name]] if name in ix else ""

        for r in rows[2:]:
            name = g(r, "Kernel Name").split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            dur = to_us(g(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]])
            rd = to_bytes(g(r, "dram__bytes_read.sum"), units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(g(r, "dram__bytes_write.sum"), units[ix["dram__bytes_write.sum"]])
            gbs = (rd + wr) / (dur * 1e-6) / 1e9 if dur > 0 else float("nan")
            tc = g(r, "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed") or g(r, "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active")
            stalls = sorted(((num(r[ix[h]]), h) for h in stall_cols if r[ix[h]]), reverse=True)[:3]
            st = ", ".join(f"{h.split('issue_stalled_')[-1].split('_per')[0].split('.')[0]} {v:.1f}" for v, h in stalls if v == v)
            out.append(f"| `{name}` | {g(r, 'launch__grid_size')} x {g(r, 'launch__block_size')} | {g(r, 'launch__registers_per_thread')} | {dur:.1f} | "
                       f"{(rd + wr) / 1e6:.1f} | {gbs:.0f} | {100 * gbs / hbm:.0f} | {num(tc):.1f} | "
                       f"{num(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {num(g(r, 'lts__t_sector_hit_rate.pct')):.0f} | {st} |")
    text = "\n".join(out) + "\n"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "ncu_summary.md"), "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
