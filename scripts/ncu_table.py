"""Markdown table of the metrics we quote from `ncu --set full` reports.

    python scripts/ncu_table.py profiles/r02_ncu_attn_shape2.ncu-rep [more.ncu-rep ...]

One row per captured kernel launch.  Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed)."""
import csv
import subprocess
import sys

COLS = [  # (header, metric-name suffix, format)
    ("time us", "gpu__time_duration.sum", "%.1f"),
    ("dram rd MB", "dram__bytes_read.sum", "%.1f"),
    ("dram wr MB", "dram__bytes_write.sum", "%.1f"),
    ("dram %pk", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"),
    ("L2 %pk", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"),
    ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "%.1f"),
    ("tensor %", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "%.1f"),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "%.1f"),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "%.1f"),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "%.1f"),
    ("regs", "launch__registers_per_thread", "%.0f"),
    ("CTA/SM (regs,smem)", None, None),
    ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active", "%.1f"),
    ("grid", "launch__grid_size", "%.0f"),
]
UNIT_SCALE = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
              "nsecond": 1e-3, "msecond": 1e3}


def find(header, suffix):
    for i, n in enumerate(header):
        if n == suffix:
            return i
    for i, n in enumerate(header):
        if n.endswith(suffix):
            return i
    return None


def main(paths):
    print("| report | kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print("| %s | (no kernels) |" % path)
            continue
        header, units = rows[0], rows[1]
        kn = find(header, "Kernel Name")
        for r in rows[2:]:
            cells = []
            for title, suffix, fmt in COLS:
                if suffix is None:
                    a = find(header, "launch__occupancy_limit_registers")
                    b = find(header, "launch__occupancy_limit_shared_mem")
                    cells.append("%s, %s" % (r[a].split(".")[0] if a is not None else "?", r[b].split(".")[0] if b is not None else "?"))
                    continue
                i = find(header, suffix)
                if i is None or r[i] == "":
                    cells.append("")
                    continue
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    cells.append(r[i])
                    continue
                if "MB" in title or "us" in title:
                    v *= UNIT_SCALE.get(units[i], 1.0)
                cells.append(fmt % v)
            name = r[kn].split("(")[0].replace("bb::", "").replace("fat::", "")
            if "<" in r[kn].split("(")[0]:
                name = r[kn].split("(")[0]
            print("| %s | `%s` | %s |" % (path.split("/")[-1].replace(".ncu-rep", ""), name, " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1:])
