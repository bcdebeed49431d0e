"""Aggregate `ncu --page source --csv --print-source sass` output by opcode (instructions executed, stall samples)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        n, s = int(r[iN]), int(r[iSamp])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op = m.group(2) if m else "?"
    agg[op] += n; samp[op] += s; tot += n; stot += s
print("total warp-instructions", tot, "samples", stot)
for op, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(stot,1):5.1f}%")
