"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: name, launches, mean us."""
import csv
import re
import sys
from collections import OrderedDict

rows = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    key = (name[:110], r["Grid Size"])
    rows.setdefault(key, []).append(float(r["Metric Value"].replace(",", "")) / 1e3)
for (name, grid), v in rows.items():
    print(f"{len(v):5d} x {sum(v) / len(v):10.2f} us  (min {min(v):9.2f})  grid {grid:>16}  {name}")
