"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel, for the last frame that starts
with the kernel named on the command line:  python scripts/launch_summary.py list.csv [first_kernel_substring]"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
first = sys.argv[2] if len(sys.argv) > 2 else None
if first:
    idx = [i for i, d in enumerate(data) if first in d[0]]
    data = data[idx[-1]:]
agg = collections.OrderedDict()
for k, v in data:
    k = k.split("(")[0].replace("void ", "")[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
out = {"total_us": tot / 1e3, "kernels": [{"name": k, "launches": c, "us": v / 1e3, "share": v / tot} for k, (c, v) in agg.items()]}
for e in out["kernels"]:
    print("%-72s %3d %9.1f us %5.1f%%" % (e["name"], e["launches"], e["us"], 100 * e["share"]))
print("total %.1f us" % out["total_us"])
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
