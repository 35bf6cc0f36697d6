"""Per-launch durations of the rowlin kernel from an ncu launch list (last training step): python tools/rowlin_list.py <csv>"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = [float(r[vi].replace(",", "")) / 1e3 for r in rows[1:] if "rowlin" in r[ki]]
last = d[-26:]
print("fwd:", " ".join("%.1f" % x for x in last[:13]))
print("bwd:", " ".join("%.1f" % x for x in last[13:]))
print("sum %.1f us" % sum(last))
