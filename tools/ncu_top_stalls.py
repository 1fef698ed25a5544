"""Top stall locations of one kernel launch from `ncu -i X.ncu-rep --page source --csv` output.
Usage: python tools/ncu_top_stalls.py src.csv [launch_index] [top_n]"""
import csv
import sys

path, which, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
with open(path) as f:
    for row in csv.reader(f):
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(row)
blk = blocks[which]
hdr = blk[0]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
rows = blk[1:]
tot = sum(int(r[ix["# Samples"]]) for r in rows)
print(f"launch {which}: {len(rows)} instructions, {tot} samples")
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]]))[:topn]
for i in sorted(order):
    r = rows[i]
    n = int(r[ix["# Samples"]])
    top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
    tops = " ".join(f"{s}:{v}" for v, s in top if v)
    print(f"{i:5d} {100*n/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} ex={r[ix['Instructions Executed']]:>9s}  {tops}")
