"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line: warp instructions,
average active threads, lane slots lost (32 x warp instructions - thread instructions).
usage: python scripts/ncu_line_summary.py src.csv [top_n] [file_substring]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else None
cur = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        # two "Source" columns: first = CUDA text, second = SASS
        src_cols = [i for i, k in enumerate(r) if k == "Source"]
        continue
    if hdr is None or cur is None or not r[0].strip().isdigit():
        continue
    try:
        w = int(float(r[hdr["Instructions Executed"]] or 0))
        t = int(float(r[hdr["Thread Instructions Executed"]] or 0))
        smp = int(float(r[hdr["# Samples"]] or 0))
    except (ValueError, IndexError):
        continue
    key = (cur.split("/")[-1], int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, r[src_cols[0]].strip()])
    a[0] += w
    a[1] += t
    a[2] += smp
tw = sum(a[0] for a in agg.values())
tt = sum(a[1] for a in agg.values())
print("warp instructions %d, thread instructions %d, avg active %.2f, lane slots lost %.1f%%" % (tw, tt, tt / tw, 100 * (1 - tt / (32.0 * tw))))
items = [(k, a) for k, a in agg.items() if a[0] and (only is None or only in k[0])]
print("top lines by lane slots lost:")
for k, a in sorted(items, key=lambda ka: -(32 * ka[1][0] - ka[1][1]))[:top_n]:
    print("  %-26s:%4d  warp %5.2f%%  avg thr %5.2f  lost %5.2f%% of all slots   %s" % (
        k[0], k[1], 100.0 * a[0] / tw, a[1] / a[0], 100.0 * (32 * a[0] - a[1]) / (32.0 * tw), a[3][:90]))
