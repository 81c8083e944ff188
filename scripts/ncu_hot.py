"""Hottest SASS instructions (by warp-stall samples) of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv; python scripts/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
first = next(i for i, h in enumerate(hdr) if h.startswith("stall_"))
stalls = [i for i in range(first, len(hdr)) if "(Not Issued)" not in hdr[i]]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which launch block of the file
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Address":
        cur = []
        blocks.append(cur)
    elif cur is not None and len(r) == len(hdr) and r[0].startswith("0x"):
        cur.append(r)
data = blocks[which]
tot = sum(int(r[si]) for r in data)
agg = {}
for r in data:
    for i in stalls:
        if r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("instructions %d, samples %d; by reason: %s" % (
    len(data), tot, ", ".join("%s %.0f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8])))
for n, r in sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:topn]:
    st = sorted(((hdr[i], int(r[i])) for i in stalls if r[i].isdigit() and int(r[i]) > 0), key=lambda x: -x[1])[:3]
    print("%5d  smp %6s  exec %7s  %-64s %s" % (n, r[si], r[ie], r[1].strip()[:64], st))
