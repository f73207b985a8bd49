"""Per-source-line view of an .ncu-rep (needs -lineinfo): executed instructions, stall samples by reason, shared-memory
wavefronts (ideal / excessive).  Usage: python tools/ncu_src.py <report.ncu-rep> [top=25]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur = ""
L = collections.defaultdict(collections.Counter)
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    src[(cur, ln)] = r[1].strip()
    for j, h in enumerate(hdr):
        if h in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive") or (
                h.startswith("stall_") and "Not Issued" not in h):
            try:
                L[(cur, ln)][h] += int(r[j])
            except (ValueError, IndexError):
                pass
tot = collections.Counter()
for c in L.values():
    tot.update(c)
S, I, W, E = tot["# Samples"], tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"]
print(f"samples {S}  warp-instructions {I}  shared wavefronts {W} (excessive {E})")
print("stall mix:", {k[6:]: round(100 * v / S, 1) for k, v in tot.most_common() if k.startswith("stall_") and v > 0.01 * S})
for title, key in (("stall samples", "# Samples"), ("shared wavefronts", "L1 Wavefronts Shared"), ("instructions", "Instructions Executed")):
    print(f"--- top lines by {title}")
    for k, c in sorted(L.items(), key=lambda kv: -kv[1][key])[:top]:
        main = max(((h[6:], v) for h, v in c.items() if h.startswith("stall_")), key=lambda t: t[1], default=("", 0))[0]
        print(f"{100*c['# Samples']/S:5.1f}% smp {100*c['Instructions Executed']/max(I,1):5.1f}% ins {100*c['L1 Wavefronts Shared']/max(W,1):5.1f}% wf "
              f"{main:12s} {k[0]}:{k[1]}  {src[k][:80]}")
