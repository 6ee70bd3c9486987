"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of the LAST forward in the file."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
seq = [(r["Kernel Name"].split("(")[0][-42:], r["Grid Size"], float(r["Metric Value"].replace(",", "")) / (1000.0 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)) for r in rows]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
tot = 0.0
for k, g, t in seq[-n:]:
    tot += t
    print(f"  {k:44s} {g:18s} {t:9.1f} us")
print(f"  sum of the last {n}: {tot:.1f} us")
