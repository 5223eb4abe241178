#!/usr/bin/env python
"""Reads an .ncu-rep (ncu --set full) here on the CPU box and prints one markdown row per profiled launch with the
metrics the roofline discussion uses; with --traffic KEY it also records dram read+write bytes per launch of the
matching kernel in profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).

  python scripts/ncu_summary.py gpurun_out/prof_attn_r2a.ncu-rep [--traffic attn_bwd@debug-8k:attn_bwd2_kernel] [--source TEXT]
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "lsu_smem_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
}


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * f.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--traffic", action="append", default=[], help="KEY:kernel-substring[:launch-index] -> profiles/ncu_traffic.json")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    launches = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "?").split("(")[0]}
        for k, name in WANT.items():
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if name in ("dram_read", "dram_write", "l2_bytes"):
                    v = to_bytes(v, u[k])
                if name == "duration":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u[k], 1.0)
                rec[name] = v
        launches.append(rec)
    print("| # | kernel | grid x block | duration | dram read | dram write | L2 bytes | tensor pipe (% elapsed, realtime) | smem wavefronts: tensor-core operands | smem wavefronts: LSU | XU (% elapsed) | DRAM % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for i, r in enumerate(launches):
        g = lambda k, f="%.1f": (f % r[k]) if k in r else "-"   # noqa: E731
        tens = r.get("tensor_pct_elapsed", r.get("tensor_pct_elapsed2"))
        print(f"| {i} | {r['kernel']} | {g('grid', '%.0f')} x {g('block', '%.0f')} | {g('duration')} us | "
              f"{r.get('dram_read', 0) / 1e6:.1f} MB | {r.get('dram_write', 0) / 1e6:.1f} MB | {r.get('l2_bytes', 0) / 1e6:.0f} MB | "
              f"{('%.1f' % tens) if tens is not None else '-'} % | {g('smem_pipe_pct')} % | {g('lsu_smem_pct')} % | {g('xu_pct')} % | "
              f"{g('dram_pct')} % | {g('regs', '%.0f')} |")
    if a.traffic:
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        try:
            tj = json.load(open(path))
        except Exception:
            tj = {}
        for spec in a.traffic:
            parts = spec.split(":")
            key, sub = parts[0], parts[1]
            sel = [r for r in launches if sub in r["kernel"]]
            if len(parts) > 2:
                sel = [sel[int(parts[2])]]
            if not sel:
                continue
            total = sum(r.get("dram_read", 0) + r.get("dram_write", 0) for r in sel)
            tj[key] = {"dram_bytes_per_launch": total, "launches_summed": [r["kernel"] for r in sel],
                       "source": a.source or os.path.basename(a.rep)}
        json.dump(tj, open(path, "w"), indent=1)
        print("updated", path)


if __name__ == "__main__":
    main()
