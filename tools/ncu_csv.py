"""Pivot of an `ncu --csv --metrics ...` log: one line per launch.  python tools/ncu_csv.py <file.csv>"""
import csv
import sys
from collections import OrderedDict

rows = OrderedDict()
for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')):
    k = (r["ID"], r["Kernel Name"].split("(")[0][-40:])
    rows.setdefault(k, {})[r["Metric Name"].split(".")[0].replace("__", ":")] = (r["Metric Value"], r["Metric Unit"])
for (i, name), m in rows.items():
    print(i, name, " ".join("%s=%s%s" % (k, v[0], v[1] if v[1] in ("ns", "%") else "") for k, v in m.items()))
