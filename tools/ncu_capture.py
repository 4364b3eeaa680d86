#!/usr/bin/env python
"""tools/ncu_capture.py -- turn an ncu metric pass over tools/profile_run.py into profiles/ncu_capture.json (the per-stage DRAM
traffic and pipe utilisation bench.py quotes in its `roofline` object) and a per-kernel table.

On the GPU box (after the same command ran clean without ncu):
    ncu --metrics <METRICS below> --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches.csv \
        python tools/profile_run.py --reps 2
Here:
    python tools/ncu_capture.py gpurun_out/launches.csv [profiles/ncu_capture.json] [--note "..."]
"""
import collections
import csv
import json
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,"
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,"
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,"
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
STAGE = (("k_pyr", "pyramid"), ("k_fast", "fast"), ("k_octree", "octree"), ("k_blur", "blur"), ("k_describe", "describe"))


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    return float(v) * m.get(unit, 1)


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        print("METRICS=" + METRICS)
        return 0
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
        d[r[ix["Metric Name"]]] = (r[ix["Metric Value"]].replace(",", ""), r[ix["Metric Unit"]])
    stages = collections.OrderedDict()
    print("%-34s %9s %9s %9s %9s %7s %6s %6s %6s %6s %6s" % ("kernel", "us", "Minstr", "dramR MB", "dramW MB", "L2 MB", "alu%", "fma%", "lsu%", "issue%", "warps%"))
    for d in launches.values():
        g = lambda k: float(d[k][0]) if k in d else float("nan")
        us = to_us(*d["gpu__time_duration.sum"])
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        l2 = to_bytes(*d["lts__t_bytes.sum"]) if "lts__t_bytes.sum" in d else float("nan")
        inst = g("smsp__inst_executed.sum")
        alu, fma, lsu = (g("sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % p) for p in ("alu", "fma", "lsu"))
        issue, warps = g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active")
        print("%-34s %9.1f %9.1f %9.1f %9.1f %7.0f %6.1f %6.1f %6.1f %6.1f %6.1f" % (d["name"][:34], us, inst / 1e6, rd / 1e6, wr / 1e6, l2 / 1e6, alu, fma, lsu, issue, warps))
        st = next((s for p, s in STAGE if d["name"].lstrip("void ").startswith(p)), None)
        if st is None:
            continue
        a = stages.setdefault(st, {"us": 0.0, "dram": 0.0, "inst": 0.0, "w": 0.0, "alu": 0.0, "fma": 0.0, "lsu": 0.0, "issue": 0.0})
        a["us"] += us; a["dram"] += rd + wr; a["inst"] += inst
        for k, v in (("alu", alu), ("fma", fma), ("lsu", lsu), ("issue", issue)):
            a[k] += v * us
        a["w"] += us
    res = {"_note": note or "per-stage DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum over all kernels of the stage) and time-weighted pipe "
                           "utilisation for one launch group of 256 frames, from an ncu metric pass over tools/profile_run.py"}
    tot = sum(a["us"] for a in stages.values())
    for st, a in stages.items():
        res[st] = {"frames_per_launch": 256, "dram_bytes_per_launch": a["dram"], "us_under_ncu": a["us"], "share_of_kernel_time": a["us"] / tot,
                   "warp_instructions": a["inst"],
                   "pipes": {"alu_pct": round(a["alu"] / a["w"], 1), "fma_pct": round(a["fma"] / a["w"], 1), "lsu_pct": round(a["lsu"] / a["w"], 1),
                             "issue_active_pct": round(a["issue"] / a["w"], 1)}}
        print("%-10s %8.1f us  %5.1f %%  %8.1f MB DRAM  %8.1f M warp-instr" % (st, a["us"], 100 * a["us"] / tot, a["dram"] / 1e6, a["inst"] / 1e6))
    if out:
        json.dump(res, open(out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
