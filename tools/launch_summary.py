#!/usr/bin/env python
"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] == "gpu__time_duration.sum":
        k = re.sub(r"\(.*", "", r[ki]).replace("csrk::", "").replace("void ", "")[:80]
        agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("| share | launches | avg µs | kernel |\n|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {100*sum(v)/tot:.2f}% | {len(v)} | {sum(v)/len(v)/1000:.2f} | `{k}` |")
