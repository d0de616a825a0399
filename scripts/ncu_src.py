"""Summarises the source page of an ncu report (per-opcode samples / executed counts, stall totals, hottest SASS lines).
usage: python scripts/ncu_src.py report.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for bl in blocks:
    if want not in bl["name"]:
        continue
    h, data = bl["rows"][0], bl["rows"][1:]
    ci = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    print(bl["name"][:70], "| samples", tot, "| warp-instr", sum(int(r[ci["Instructions Executed"]]) for r in data))
    op, ex = Counter(), Counter()
    for r in data:
        w = r[ci["Source"]].split()
        o = w[1] if w[0].startswith("@") else w[0]
        op[o] += int(r[ci["# Samples"]])
        ex[o] += int(r[ci["Instructions Executed"]])
    print("  samples by opcode:", op.most_common(12))
    print("  executed by opcode:", ex.most_common(16))
    st = Counter()
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            st[k] = sum(int(r[ci[k]]) for r in data)
    print("  stalls:", st.most_common(9))
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:14]:
        print("   ", r[ci["# Samples"]], r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:64], "| long_sb", r[ci["stall_long_sb"]],
              "wait", r[ci["stall_wait"]], "barrier", r[ci["stall_barrier"]], "math", r[ci["stall_math"]], "short", r[ci["stall_short_sb"]])
