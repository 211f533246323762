"""Summarise an ncu launch-list report (gpu__time_duration.sum per launch) into per-kernel totals.
usage: python tools/ncu_summary.py gpurun_out/launches.ncu-rep [out.md]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ki, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
unit = rows[1][ti]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[ki]
    for a, b in (("vd::tc::", ""), ("vd::(anonymous namespace)::", ""), ("(CUtensorMap_st, CUtensorMap_st, ", "("), ("void ", "")):
        name = name.replace(a, b)
    name = name.split("(")[0]
    grid = r[gi].replace(" ", "") if gi is not None else ""
    big = ""
    if name.startswith("k_tc_gemm<") and gi is not None:
        big = " [grid %s]" % grid.split(",")[0].strip("(")
    key = name + big
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ti]) * scale
tot = sum(v[1] for v in agg.values())
lines = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
lines.append("| **total** | %d | %.1f | | |" % (sum(v[0] for v in agg.values()), tot))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
