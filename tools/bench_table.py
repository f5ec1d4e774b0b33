"""Rewrites the headline table of profiles/README.md from profiles/r02_bench_n1.json (the driver-format line of `python bench.py`)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read())
R, F, W = d["routines"], d["routines_f64"], d["routines_l2_walk"]


def row(name, key, dbl):
    r, f = R[key], F.get(dbl)
    return f"| {name} | {r['ms']:.4f} | {r['GBps']:.1f} | {r['pct_of_8TBs_per_gpu']:.1f} | {r['frac_of_measured_peak']:.3f} | {dbl + ' ' + format(f['GBps'], '.1f') if f else ''} |"


step_ms = d["ms_per_step"] / d["config"]["passes_per_step"]
lines = ["| routine | ms | GB/s | % of 8 TB/s | of measured | Double instance GB/s |", "|---|---|---|---|---|---|",
         f"| step = sdot+saxpy+snrm2 (the `value`, per pass) | {step_ms:.4f} | {d['value']:.1f} | {100 * d['value'] / 8000:.1f} | {d['value'] / 6558.1:.3f} | |"]
for name, key, dbl in [("sdot", "sdot", "ddot"), ("saxpy", "saxpy", "daxpy"), ("snrm2", "snrm2", "dnrm2"), ("sasum", "sasum", "dasum"),
                       ("isamax", "isamax", "idamax"), ("sscal", "sscal", "dscal"), ("scopy", "scopy", "dcopy"), ("sswap", "sswap", "dswap"),
                       ("srot", "srot", "drot")]:
    lines.append(row(name, key, dbl))
a, b, c = R["srotm"], R["srotm_FLAGNEG1"], R["srotm_FLAG1"]
lines.append(f"| srotm FLAG0 / FLAGNEG1 / FLAG1 | {a['ms']:.4f} / {b['ms']:.4f} / {c['ms']:.4f} | {a['GBps']:.0f} / {b['GBps']:.0f} / {c['GBps']:.0f} | "
             f"{a['pct_of_8TBs_per_gpu']:.1f} / {b['pct_of_8TBs_per_gpu']:.1f} / {c['pct_of_8TBs_per_gpu']:.1f} | {a['frac_of_measured_peak']:.3f} | drotm {F['drotm']['GBps']:.1f} |")
r = R["sdsdot"]
lines.append(f"| sdsdot | {r['ms']:.4f} | {r['GBps']:.1f} | {r['pct_of_8TBs_per_gpu']:.1f} | {r['frac_of_measured_peak']:.3f} | |")
p = os.path.join(ROOT, "profiles", "README.md")
s = open(p).read()
i = s.index("| routine | ms | GB/s | % of 8 TB/s | of measured | Double instance GB/s |")
s = s[:i] + "\n".join(lines) + s[s.index("\n\n", i):]
ev = d["routines_everyday"]
s = re.sub(r"sdot \d+, snrm2 \d+, sasum \d+, isamax \d+\nGB/s", f"sdot {ev['sdot']['GBps']:.0f}, snrm2 {ev['snrm2']['GBps']:.0f}, sasum {ev['sasum']['GBps']:.0f}, isamax {ev['isamax']['GBps']:.0f}\nGB/s", s, count=1)
s = re.sub(r"`roofline` \(saxpy in step context\): [\d.]+ / 6558.1 = [\d.]+", f"`roofline` (saxpy in step context): {d['roofline']['achieved']:.1f} / 6558.1 = {d['roofline']['frac']:.3f}", s, count=1)
s = re.sub(r"give saxpy [\d.]+, sscal [\d.]+, scopy [\d.]+, sswap [\d.]+, srot [\d.]+ %",
           f"give saxpy {W['saxpy']['pct_of_8TBs_per_gpu']}, sscal {W['sscal']['pct_of_8TBs_per_gpu']}, scopy {W['scopy']['pct_of_8TBs_per_gpu']}, sswap {W['sswap']['pct_of_8TBs_per_gpu']}, srot {W['srot']['pct_of_8TBs_per_gpu']} %", s, count=1)
open(p, "w").write(s)
print("\n".join(lines))
