"""Summaries of ncu outputs for profiles/ (run on the CPU box; needs the `ncu` CLI only to read reports).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/<name>.json
  python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep            > profiles/<name>.json
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct")


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)  # -> microseconds
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    out = {"source": path, "note": "ncu launch list, gpu__time_duration.sum per launch: cold-cache, serialised -- compare SHARES",
           "total_us": total, "launches": sum(c for c, _ in agg.values()), "kernels": []}
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out["kernels"].append({"kernel": k, "launches": c, "total_us": round(t, 1), "mean_us": round(t / c, 2),
                               "share": round(t / total, 4)})
    return out


def report(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = {"source": path, "kernels": []}
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for i, hname in enumerate(hdr):
            if hname in KEEP and i < len(r):
                try:
                    d[hname] = float(r[i].replace(",", ""))
                except ValueError:
                    d[hname] = r[i]
                d[hname + " unit"] = units[i]
        out["kernels"].append(d)
    return out


if __name__ == "__main__":
    fn = {"launches": launches, "report": report}[sys.argv[1]]
    print(json.dumps(fn(sys.argv[2]), indent=1))
