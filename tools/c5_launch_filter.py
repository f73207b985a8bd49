"""Keep the launches of ONE c5 evaluation (the second of the two that tools/run_c5_once.py runs) from a full ncu launch list:
python tools/c5_launch_filter.py gpurun_out/f_launches_c5_all.csv profiles/r1_launches_c5_large.csv"""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = open(src).read().splitlines()
pre = [l for l in lines if l.startswith("==")]
body = [l for l in lines if not l.startswith("==")]
hdr, rows = body[0], body[1:]
fin = [i for i, l in enumerate(rows) if "large_finish_kernel" in l]
sec = rows[fin[0] + 1:fin[1] + 1]
sec = sec[next(i for i, l in enumerate(sec) if "large_" in l):]
open(dst, "w").write("\n".join(pre[:1] + [hdr] + sec) + "\n")
r = list(csv.reader([hdr] + sec))
ki, vi = r[0].index("Kernel Name"), r[0].index("Metric Value")
t, n = collections.Counter(), collections.Counter()
for x in r[1:]:
    k = x[ki].split("(")[0]
    t[k] += float(x[vi])
    n[k] += 1
tot = sum(t.values())
print(f"{len(sec)} launches, {tot / 1e6:.3f} ms (serialised, cold)")
for k, v in t.most_common(6):
    print(f"  {100 * v / tot:5.1f}%  {n[k]:4d} x {k[:60]}  avg {v / n[k] / 1e3:.1f} us")
