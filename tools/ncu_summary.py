"""Print the headline metrics of an .ncu-rep (raw page) and the top stall lines of the source page."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit",
        "smsp__issue_active.avg.pct", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "gpu__dram_throughput", "sm__inst_executed_pipe_uniform", "sm__inst_executed_pipe_tc", "dram__cycles_active"]
for h, u, v in zip(hdr, units, vals):
    if any(k in h for k in keys) and "per_second" not in h and ".max" not in h and ".min" not in h:
        print(f"{h} [{u}] = {v}")
print("--- stall reasons (per issue active)")
for h, u, v in zip(hdr, units, vals):
    if "warps_issue_stalled" in h and "per_issue_active" in h:
        try:
            if float(v) > 0.05: print(f"  {h.split('stalled_')[1].split('_per_issue')[0]}: {float(v):.2f}")
        except ValueError: pass
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # find header
    for i, r in enumerate(rows):
        if "Source" in r and any("Sampling" in c for c in r):
            hdr = r; body = rows[i + 1:]; break
    else:
        print("no source page"); sys.exit(0)
    si = hdr.index("Source")
    col = [j for j, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
    j = col[0]
    tot = 0; items = []
    for r in body:
        try: n = int(r[j])
        except (ValueError, IndexError): continue
        tot += n; items.append((n, r[si][:130]))
    items.sort(reverse=True)
    print(f"--- top source lines by stall samples (total {tot})")
    for n, s in items[:int(sys.argv[2])]:
        print(f"  {100.0 * n / max(tot, 1):5.1f}%  {s}")
