"""Per-instruction issue-stall fields of a kernel's SASS (bits 105..108 of the 128-bit encoding) + a static estimate of
the single-warp issue time of an address range:  python tools/sass_stalls.py lib.so kernel-regex [lo hi]"""
import re, subprocess, sys
lib, pat = sys.argv[1], re.compile(sys.argv[2])
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on, cur = False, None
rows = []
for i, line in enumerate(out):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout
        on = bool(pat.search(name))
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m and cur:
        hi64 = int(m.group(1), 16)
        stall = (hi64 >> (105 - 64)) & 0xf
        yld = (hi64 >> (109 - 64)) & 1
        wbar = (hi64 >> (110 - 64)) & 7
        rbar = (hi64 >> (113 - 64)) & 7
        wmask = (hi64 >> (116 - 64)) & 0x3f
        rows.append((cur[0], cur[1], stall, yld, wbar, rbar, wmask))
        cur = None
sel = [r for r in rows if lo <= r[0] < hi]
tot = sum(r[2] for r in sel)
print(f"{len(sel)} instructions, sum of stall fields {tot} cycles ({tot/max(len(sel),1):.2f} per instruction)")
if len(sys.argv) > 5:
    for r in sel:
        print(f"{r[0]:05x} st={r[2]:2d} y={r[3]} wb={r[4]} rb={r[5]} wm={r[6]:02x}  {r[1][:100]}")
else:
    import collections
    h = collections.Counter(r[2] for r in sel)
    print("stall histogram:", sorted(h.items()))
    big = collections.Counter()
    for r in sel:
        big[r[1].split()[0] if not r[1].startswith('@') else r[1].split()[1]] += r[2]
    print("stall cycles by opcode:", big.most_common(12))
