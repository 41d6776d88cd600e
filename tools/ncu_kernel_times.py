"""Prints kernel name, duration of every launch in an `ncu --metrics gpu__time_duration.sum --csv` log (development aid)."""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
for r in rows[1:]:
    if len(r) > vi: print(r[ki][:60].replace("<unnamed>::", ""), r[vi], r[ui])
