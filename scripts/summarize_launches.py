"""Summarises an ncu launch list (gpu__time_duration.sum per launch, --csv) by kernel: launches, total time, share.
usage: python scripts/summarize_launches.py launches.csv [steps]   (steps: to print per-step figures)"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4]
    m = re.match(r"(?:void )?((?:\w+::)*\w+)", name)
    key = m.group(1) if m else name[:50]
    if "gemm_tc_kernel" in name:
        mm = re.search(r"gemm_tc_kernel<\(int\)(\d), \(int\)(\d)>", name)
        key = "bb::gemm_tc_kernel<ctas=%s,epi=%s>" % (mm.group(1), mm.group(2)) if mm else key
    tot[key][0] += 1
    tot[key][1] += float(r[14]) * 1e-6
total = sum(v[1] for v in tot.values())
print("total %.2f ms over %d launches%s" % (total, len(rows), "  (%.2f ms, %.0f launches per step)" % (total / steps, len(rows) / steps) if steps else ""))
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-58s %6d %9.2f ms %5.1f%%  avg %7.1f us" % (k[:58], n, ms, 100 * ms / total, ms / n * 1e3))
