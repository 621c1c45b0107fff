"""Top stall-sample instructions of an ncu --set full --import-source on capture (SASS view):
   ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
iS, iSamp, iEx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) > iEx and r[iSamp].isdigit()]
tot = sum(int(r[iSamp]) for r in body)
print("total samples", tot, "instructions", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][iSamp]))[:top]
for i in sorted(idx):
    r = body[i]
    prev = body[i - 1][iS].strip() if i > 0 else ''
    print(f"{i:5d} {int(r[iSamp]):7d} {100*int(r[iSamp])/tot:5.1f}%  ex={r[iEx]:>9}  {r[iS].strip()[:90]}   <- prev: {prev[:50]}")
