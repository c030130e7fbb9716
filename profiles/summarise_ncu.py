#!/usr/bin/env python
"""Summarise `ncu --set full` reports into a markdown table + profiles/ncu_traffic.json (read by bench.py).

    python profiles/summarise_ncu.py profiles/r02_ops.ncu-rep profiles/r02_conv_tcp.ncu-rep > profiles/r02_ncu_summary.md

Per kernel (last captured launch of each name): duration, DRAM read / write bytes, achieved DRAM GB/s and its fraction of
the measured copy peak (MEASURED_PEAKS.json), tensor-pipe active %, issue-slot utilisation, registers."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
WANT = {"dur_us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "second": 1e6, "nsecond": 1e-3}


def rows_of(rep):
    # a .ncu-rep report, or the `ncu -i rep --page raw --csv` dump of one (kept when the report itself is too large)
    out = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        rec = {"name": r[col["Kernel Name"]]}
        for k, m in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                rec[k] = float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
        yield rec


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    return n.split("(")[0]


def main(reps):
    last, seen, owner = {}, {}, {}
    for rep in reps:
        for rec in rows_of(rep):
            rec["rep"] = os.path.basename(rep)
            k = short(rec["name"])
            if owner.get(k, rep) != rep:                 # a later report re-captured this kernel: it replaces the older rows
                for old in [key for key in last if key.rsplit(" #", 1)[0] == k]:
                    del last[old]
                seen.pop(k, None)
            owner[k] = rep
            seen[k] = seen.get(k, -1) + 1
            last["%s #%d" % (k, seen[k])] = rec          # every captured launch, numbered per kernel name
    print("| kernel | duration | DRAM read + write | DRAM GB/s | of measured copy peak (%.0f GB/s) | tensor pipe | issue active | regs | report |" % HBM)
    print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
    traffic = {}
    for k, r in sorted(last.items()):
        if "dur_us" not in r:
            continue
        byt = r.get("rd", 0) + r.get("wr", 0)
        gbs = byt / r["dur_us"] / 1e3
        print("| `%s` | %.1f us | %.1f + %.1f MB | %.0f | %.2f | %s | %.0f %% | %d | %s |" % (
            k, r["dur_us"], r.get("rd", 0) / 1e6, r.get("wr", 0) / 1e6, gbs, gbs / HBM,
            ("%.0f %%" % r["tensor"]) if r.get("tensor", 0) > 0.5 else "-", r.get("issue", 0), int(r.get("regs", 0)), r["rep"]))
        traffic[k] = {"dram_bytes": byt, "duration_us": r["dur_us"], "source": "profiles/" + r["rep"]}
    alias = {"conv_tc2_res3_conv1": "conv_tcp_kernel<208, 1> #0", "att_general_fwd": "att_general_fwd_tc_kernel<20> #1"}
    for a, k in alias.items():
        if k not in traffic and k.endswith("#1") and k[:-1] + "0" in traffic:
            k = k[:-1] + "0"
        if k in traffic:
            traffic[a] = traffic[k]
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
