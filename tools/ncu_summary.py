"""Key numbers of every launch in an ncu report (--set full), as a markdown table.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time us", "time"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
        ("smsp__inst_executed.sum", "warp inst M", 1e-6), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
        ("dram__bytes_read.sum", "dram rd MB", None), ("dram__bytes_write.sum", "dram wr MB", None),
        ("lts__t_bytes.sum", "L2 MB", None), ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1)]


def to_mb(v, unit):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, None)
    return None if f is None else v * f


for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(k), n, sc) for k, n, sc in KEYS if k in hdr]
    print(f"### {rep.split('/')[-1]}\n")
    print("| kernel | " + " | ".join(n for _, n, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    ik = hdr.index("Kernel Name")
    for r in data:
        cells = []
        for i, n, sc in cols:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                cells.append(r[i]); continue
            if sc == "time":
                f = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(units[i], 1.0)
                cells.append(f"{v * f:.1f}")
            elif sc is None:
                mb = to_mb(v, units[i])
                cells.append(f"{mb:.2f}" if mb is not None else f"{v} {units[i]}")
            else:
                cells.append(f"{v * sc:.1f}")
        print("| " + r[ik].split("(")[0][:48] + " | " + " | ".join(cells) + " |")
    print()
