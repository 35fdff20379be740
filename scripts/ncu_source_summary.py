"""Summarise `ncu -i X.ncu-rep --page source --csv` output: opcode mix weighted by executed count, stall mix,
and the hottest SASS lines.  usage: python scripts/ncu_source_summary.py src.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
# ncu prints the table of a launch twice (and one pair per selected launch): keep the first table only
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Address":
        break
    if len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
        data.append(r)
ix = {k: i for i, k in enumerate(hdr)}
N = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(N(r, "Instructions Executed") for r in data)
thr = sum(N(r, "Thread Instructions Executed") for r in data)
print("kernel:", rows[0][1] if rows[0] else "?")
print("first selected launch: SASS lines %d, warp instructions %d, thread instructions %d, avg active threads %.2f" % (
    len(data), tot, thr, thr / max(tot, 1)))
hist = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src
    hist[op] += N(r, "Instructions Executed")
print("opcode mix (warp instructions):")
for op, n in hist.most_common(30):
    print("  %-10s %12d %5.1f%%" % (op, n, 100.0 * n / tot))
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
s = collections.Counter()
for r in data:
    for k in st:
        s[k] += N(r, k)
tt = sum(s.values())
print("stall samples:", {k: round(100.0 * v / tt, 1) for k, v in s.most_common(10)})
print("hottest lines by samples:")
for r in sorted(data, key=lambda r: -N(r, "# Samples"))[:top_n]:
    print("  %6d samp %10d exec  %s" % (N(r, "# Samples"), N(r, "Instructions Executed"), r[ix["Source"]].strip()))
