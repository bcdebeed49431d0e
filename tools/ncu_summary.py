"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/."""
import csv, subprocess, sys, io, re, collections
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (one row per captured launch)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        for k in keys:
            if k in d: f.write(f"{k:75s} {d[k]} {u.get(k,'')}\n")
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) > 0.1: f.write(f"stall {k.split('stalled_')[1].split('_per')[0]:28s} {v}\n")
                except ValueError: pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src))); h = rows[1]
    iS, iN = h.index("Source"), h.index("Instructions Executed")
    agg, tot = collections.Counter(), 0
    for r in rows[2:]:
        try: n = int(r[iN])
        except (ValueError, IndexError): continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS]); agg[m.group(2) if m else "?"] += n; tot += n
    f.write(f"\n# SASS opcode mix (warp-instructions executed, total {tot})\n")
    for op, n in agg.most_common(16): f.write(f"{op:10s} {n:12d} {100*n/tot:5.1f}%\n")
print(open(out).read())
