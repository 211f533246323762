"""Key metrics of an `ncu --set full` report, one row per profiled launch.
usage: python tools/ncu_full_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
cols = [(c, n) for c, n in cols if c in hdr]
lines = ["| " + " | ".join("%s%s" % (n, (" [%s]" % units[hdr.index(c)]) if units[hdr.index(c)] and n not in ("kernel",) and "%" not in n else "") for c, n in cols) + " |",
         "|" + "---|" * len(cols)]
for r in rows[2:]:
    vals = []
    for c, n in cols:
        v = r[hdr.index(c)]
        if n == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "") + "`"
        else:
            try:
                v = "%.1f" % float(v.replace(",", ""))
            except ValueError:
                pass
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
