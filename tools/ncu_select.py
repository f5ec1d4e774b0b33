"""Reduce `ncu -i <rep> --page raw --csv` to the columns the roofline discussion uses and refresh
profiles/traffic.json (DRAM bytes per launch of each headline kernel).
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_select.py /tmp/raw.csv profiles/r01_ncu_full_selected.csv profiles/traffic.json
"""
import csv
import json
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
ROUTINE = {"AxpyF": "saxpy", "DotOp": "sdot", "Nrm2Op": "snrm2", "IamaxOp": "isamax", "SwapF": "sswap", "CopyF": "scopy", "AsumOp": "sasum", "Rotm0F": "srotm",
           "ScalF": "sscal", "RotF": "srot"}

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
cols = [hdr.index(k) for k in KEEP if k in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[c] for c in cols]); w.writerow([units[c] for c in cols])
    for r in body:
        w.writerow([r[c] for c in cols])
if len(sys.argv) > 3:
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    out = {"source": f"{sys.argv[2]}: dram__bytes_read.sum + dram__bytes_write.sum per launch at n=2^28 (ncu --set full --clock-control none, one launch each)"}
    for r in body:
        for tag, name in ROUTINE.items():
            if tag + "<" in r[ik] and name not in out:
                out[name] = int(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
                out[name + "_kernel"] = r[ik].split("(")[0].replace("void ", "").strip()
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out, indent=1))
