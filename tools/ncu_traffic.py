#!/usr/bin/env python
"""Turn the captures of tools/ncu_traffic.sh into profiles/r2_traffic.json (what bench.py's roofline.traffic reads) and the
metric tables profiles/r2_ncu_<name>_summary.csv.
usage: python tools/ncu_traffic.py gpurun_out/prof_headline.ncu-rep [gpurun_out/prof_damped.ncu-rep]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def num(m, name):
    unit, val = m[name]
    v = float(val.replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}.get(unit, 1.0)


def main():
    with open(os.path.join(ROOT, "gpurun_out", "traffic_source_hash.txt")) as f:
        h = f.read().strip().splitlines()[-1]
    caps = []
    for rep in sys.argv[1:]:
        damped = "damped" in os.path.basename(rep)
        m = raw(rep)
        name = "damped" if damped else "dataflow"
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                        os.path.join(ROOT, "profiles", "r2_ncu_%s_summary.csv" % name)], check=True)
        rd, wr = num(m, "dram__bytes_read.sum"), num(m, "dram__bytes_write.sum")
        tets, sub = 998250, 50
        caps.append({"kernel": m["Kernel Name"][1].replace("void ", "").split("(")[0], "cells": 55, "precision": "exact", "energy": "yeohskinfast",
                     "damped": damped, "substeps_per_launch": sub, "tets": tets, "dram_bytes_read": rd, "dram_bytes_write": wr,
                     "dram_bytes_per_launch": rd + wr, "dram_bytes_per_element_substep": (rd + wr) / (tets * sub),
                     "l2_sectors_from_l1": float(m["lts__t_sectors_srcunit_tex.sum"][1].replace(",", "")),
                     "l2_hit_rate_pct": float(m["lts__t_sector_hit_rate.pct"][1]), "gpu_time_ms": num(m, "gpu__time_duration.sum") if m["gpu__time_duration.sum"][0] == "ms" else None,
                     "source_hash": h, "source": "profiles/r2_ncu_%s_summary.csv" % name})
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump({"captures": caps}, f, indent=1)
    print(json.dumps(caps, indent=1))


if __name__ == "__main__":
    main()
