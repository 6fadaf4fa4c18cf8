#!/usr/bin/env python
"""Text exports of an Nsight Compute report, so that the evidence in profiles/ is readable without the
binary .ncu-rep (which stays out of git):

  python tools/ncu_export.py profiles/x.ncu-rep [--kernel k_msm_accumulate] [--json profiles/x.json]

writes profiles/x.raw.csv (ncu --page raw --csv, every metric of every captured launch) and a small JSON
with the numbers DESIGN.md / bench.py quote for the chosen kernel's launches."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

KEYS = {
    "duration_ns": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "fmaheavy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "smem_bank_conflict_wavefronts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="substring of the kernel name to summarise (default: every launch)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    base = a.report[: -len(".ncu-rep")] if a.report.endswith(".ncu-rep") else a.report
    open(base + ".raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}

    def find(metric):
        for name, i in col.items():
            if name == metric or name.endswith("." + metric):
                return i
        return None

    out = []
    for r in data:
        name = r[col["Kernel Name"]]
        if a.kernel and a.kernel not in name:
            continue
        rec = {"kernel": name, "launch": f"grid {r[col['Grid Size']]} block {r[col['Block Size']]}", "id": r[col["ID"]]}
        for key, metric in KEYS.items():
            i = find(metric)
            if i is None or r[i] == "":
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if key in ("duration_ns", "dram_bytes_read", "dram_bytes_write"):
                v *= UNIT_SCALE.get(units[i], 1)
            rec[key] = v
        out.append(rec)
    if a.json:
        doc = out[0] if len(out) == 1 else {"launches": out}
        doc["source"] = os.path.basename(a.report) + " (ncu --set full --clock-control none); full metric list in " + os.path.basename(base) + ".raw.csv"
        json.dump(doc, open(a.json, "w"), indent=1)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
