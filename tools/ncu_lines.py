"""Aggregate warp-stall samples and executed instructions per CUDA source line from an .ncu-rep (needs -lineinfo)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); cur_file = ""; hdr = None; tot = 0; tot_inst = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; js = hdr.index("# Samples"); ji = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    try: n = int(r[js]); ins = int(r[ji])
    except (ValueError, IndexError): continue
    if r[0] != "":   # a CUDA source line row carries the aggregate of its SASS
        key = (cur_file, r[0]); src = r[1]
        if key not in agg: agg[key] = [0, 0, src]
        agg[key][0] += n; agg[key][1] += ins; tot += n; tot_inst += ins
items = sorted(agg.items(), key=lambda kv: -kv[1][0])
print(f"total samples {tot}, warp-instructions {tot_inst}")
for (f, ln), (n, ins, src) in items[:top]:
    print(f"{100.0*n/max(tot,1):5.1f}% smp {100.0*ins/max(tot_inst,1):5.1f}% inst  {f}:{ln}  {src.strip()[:100]}")
