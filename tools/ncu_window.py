"""SASS window of one launch from `ncu --page source --csv`: python tools/ncu_window.py src.csv launch lo hi"""
import csv, sys
path, which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
blocks, cur = [], None
with open(path) as f:
    for row in csv.reader(f):
        if row and row[0] == "Kernel Name":
            cur = []; blocks.append(cur)
        elif cur is not None:
            cur.append(row)
blk = blocks[which]; hdr = blk[0]; ix = {h: i for i, h in enumerate(hdr)}
rows = blk[1:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in range(lo, min(hi, len(rows))):
    r = rows[i]
    top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    tops = " ".join(f"{s}:{v}" for v, s in top if v)
    print(f"{i:5d} {int(r[ix['# Samples']]):5d} {r[ix['Source']].strip()[:90]:90s} ex={r[ix['Instructions Executed']]:>8s} {tops}")
