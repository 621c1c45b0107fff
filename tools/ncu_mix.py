"""Dynamic opcode mix (executed warp-instructions and stall samples per opcode) of each kernel in an ncu source-page CSV."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
kern, hdr, cur = None, None, None
out = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern = r[1][:70]; cur = {'ex': collections.Counter(), 'st': collections.Counter()}; out.append((kern, cur)); continue
    if r and r[0] == 'Address':
        hdr = r; iS, iSamp, iEx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'); continue
    if cur is None or len(r) <= iEx or not r[iEx].isdigit():
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    cur['ex'][op] += int(r[iEx]); cur['st'][op] += int(r[iSamp])
for kern, c in out:
    te, ts = sum(c['ex'].values()), sum(c['st'].values())
    print(f"== {kern}: {te} warp-instructions, {ts} samples")
    for op, n in c['ex'].most_common(22):
        print(f"   {op:10s} {n:9d} {100*n/te:5.1f}%   samples {100*c['st'][op]/max(ts,1):5.1f}%")
