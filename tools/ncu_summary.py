"""Summarise .ncu-rep captures (ncu --set full) into one markdown table for profiles/: per kernel launch the duration,
tensor-pipe activity, DRAM bytes (read + write) and the achieved HBM GB/s against MEASURED_PEAKS.json.

  python tools/ncu_summary.py gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ... > profiles/r02_ncu_top_kernels.md"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    if len(rd) < 3:
        return []
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rd[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")}
        for w in WANT:
            if w in idx:
                try:
                    d[w] = float(r[idx[w]].replace(",", "")) * UNIT.get(units[idx[w]], 1.0)
                except ValueError:
                    pass
        res.append(d)
    return res


def main():
    try:
        peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    print(f"| capture | kernel | time (us) | tensor pipe % (active) | DRAM read+write (MB) | achieved GB/s (of {peak:.0f}) | issue active % | warps active % | regs |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
    for path in sys.argv[1:]:
        for d in rows_of(path):
            t = d.get("gpu__time_duration.sum", 0.0)
            by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            gbs = by / t / 1e9 if t > 0 else 0.0
            print(f"| {os.path.basename(path).replace('.ncu-rep', '')} | `{d['kernel'][:70]}` | {t * 1e6:.1f} | "
                  f"{'%.1f' % tp if tp is not None else 'n/a'} | {by / 1e6:.1f} | {gbs:.0f} ({gbs / peak:.2f}) | "
                  f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.0f} | "
                  f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.0f} | {int(d.get('launch__registers_per_thread', 0))} |")


if __name__ == "__main__":
    main()
