"""Joins tools/ncu_more.py's manifest with the ncu CSV of the same run (by launch order) -> a markdown table and a JSON:
    python tools/ncu_join.py gpurun_out/more_manifest.csv gpurun_out/more.csv profiles/r02_ncu_more
"""
import csv
import json
import sys

manifest, ncu_csv, out = sys.argv[1:4]
rows = list(csv.DictReader(l for l in open(manifest) if not l.startswith("#")))
launches = {}
order = []
with open(ncu_csv) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in launches:
        launches[i] = {"kernel": r["Kernel Name"]}
        order.append(i)
    try:
        launches[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        pass
assert len(order) == len(rows), (len(order), len(rows))
table, js = [], {}
hdr = "| call | kernel | time us | algorithmic MB | DRAM read MB | DRAM write MB | DRAM / algorithmic | DRAM / 32-byte-sector estimate | algorithmic GB/s | DRAM GB/s | DRAM active % | read-active % | write-active % |"
table.append(hdr)
table.append("|" + "---|" * (hdr.count("|") - 1))
for row, i in zip(rows, order):
    m = launches[i]
    t_us = m.get("gpu__time_duration.sum", 0) / 1e3
    rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
    alg, sect = float(row["algorithmic_bytes"]), float(row["sector_bytes"])
    el = m.get("dram__cycles_elapsed.sum", 0) or 1
    act, ar, aw = (100 * m.get(k, 0) / el for k in ("dram__cycles_active.sum", "dram__cycles_active_read.sum", "dram__cycles_active_write.sum"))
    short = m["kernel"].split("(")[0].replace("lbk::", "")[:60]
    table.append(f"| {row['label']} | {short} | {t_us:.1f} | {alg / 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / alg:.3f} | {(rd + wr) / sect:.3f} | "
                 f"{alg / t_us / 1e3:.0f} | {(rd + wr) / t_us / 1e3:.0f} | {act:.1f} | {ar:.1f} | {aw:.1f} |")
    js[row["label"]] = {"kernel": m["kernel"], "time_us": round(t_us, 2), "algorithmic_bytes": alg, "sector_estimate_bytes": sect,
                        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_over_algorithmic": round((rd + wr) / alg, 4),
                        "traffic_over_sector_estimate": round((rd + wr) / sect, 4), "dram_cycles_active_pct": round(act, 2),
                        "dram_cycles_active_read_pct": round(ar, 2), "dram_cycles_active_write_pct": round(aw, 2)}
open(out + ".md", "w").write("\n".join(table) + "\n")
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(table))
