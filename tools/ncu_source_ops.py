"""Per-opcode executed-instruction and stall-sample totals from an `ncu --page source --csv` export (SASS view)."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); smp = collections.Counter(); stalls = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex] or not r[iex].isdigit():  # (a multi-kernel export repeats its header rows)
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
    op = src.split()[0].rstrip(";") if src else "?"
    n = int(r[iex]); ops[op] += n; tot += n; smp[op] += int(r[ismp] or 0)
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
print("total warp instructions", tot)
for op, n in ops.most_common(25):
    print(f"{n:14d} {100*n/tot:5.1f}%  samples {smp[op]:8d}  {op}")
st = sum(stalls.values())
print("stall samples:", ", ".join(f"{k[6:]} {100*v/st:.1f}%" for k, v in stalls.most_common(8)))
