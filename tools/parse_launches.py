"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step, kernel by kernel."""
import csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
seq = [(r[ki].split("(")[0].replace("mmlrec::", "")[:40], float(r[vi].replace(",", "")), r[gi]) for r in data]
starts = [i for i, s in enumerate(seq) if "hyper_advance" in s[0]]
a, b = starts[0], starts[1]
tot, agg = 0.0, {}
for name, ns, grid in seq[a:b]:
    print(f"{name:40s} {ns / 1000:9.2f} us  grid {grid}")
    tot += ns
    agg[name] = agg.get(name, 0.0) + ns
print(f"total kernel time per step: {tot / 1000:.1f} us over {b - a} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"   {k:40s} {v / 1000:9.1f} us  {100 * v / tot:5.1f}%")
