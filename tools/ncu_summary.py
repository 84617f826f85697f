#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the convolution kernels: per-launch table + totals (json + markdown).
Usage: python tools/ncu_summary.py gpurun_out/conv_v3.ncu-rep profiles/r01_ncu_conv_v3 [--title "..."]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * m
    return v


def to_us(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    launches = []
    for r in data:
        d = {}
        for k, i in idx.items():
            d[k] = (r[i], units[i])
        launches.append(d)
    tot_r = sum(to_bytes(*l["dram__bytes_read.sum"]) for l in launches)
    tot_w = sum(to_bytes(*l["dram__bytes_write.sum"]) for l in launches)
    tot_t = sum(to_us(*l["gpu__time_duration.sum"]) for l in launches)
    summ = {"report": rep, "launches": len(launches), "dram_bytes_read": tot_r, "dram_bytes_write": tot_w,
            "dram_bytes_total": tot_r + tot_w, "time_us_total_under_ncu": tot_t,
            "tensor_pipe_pct_time_weighted": sum(float(l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]) *
                                                 to_us(*l["gpu__time_duration.sum"]) for l in launches) / max(tot_t, 1e-9)}
    summ["per_launch"] = [{"kernel": l.get("Kernel Name", ("", ""))[0], "dram__bytes_read.sum": to_bytes(*l["dram__bytes_read.sum"]),
                           "dram__bytes_write.sum": to_bytes(*l["dram__bytes_write.sum"]), "time_us": to_us(*l["gpu__time_duration.sum"]),
                           "tensor_pipe_pct": float(l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                           "lts_hit_pct": float(l["lts__t_sector_hit_rate.pct"][0]) if "lts__t_sector_hit_rate.pct" in l else None}
                          for l in launches]
    json.dump(summ, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"{len(launches)} launches; DRAM read {tot_r / 1e9:.3f} GB, write {tot_w / 1e9:.3f} GB; "
                f"time under ncu {tot_t / 1e3:.3f} ms; time-weighted tensor pipe active {summ['tensor_pipe_pct_time_weighted']:.1f} %\n\n")
        cols = [k for k in KEYS if k in idx and k != "Kernel Name"]
        f.write("| # | " + " | ".join(c.replace("_", "\\_") for c in cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n")
        for i, l in enumerate(launches):
            f.write(f"| {i} | " + " | ".join(f"{l[c][0]} {l[c][1]}" for c in cols) + " |\n")
    print(json.dumps(summ))


if __name__ == "__main__":
    main()
